"""TEST INFRASTRUCTURE: builds tests/emu/_emu_cuda.so — the product's C-ABI library
(nhd_b200/csrc/nhd_api.cu + nhd_kernels.cuh + nhd_core.cuh + nhd_ingest.cpp) compiled with g++ on top
of the CPU emulation of CUDA in tests/emu/cuda_emu.h, so the real kernels can run where there is no GPU.

The sources are used as they are except for three mechanical rewrites g++ needs (done on copies under
tests/emu/_gen/, never on the product files):
  * ``kernel<<<grid, block, smem, stream>>>(args);``  ->  ``EMU_LAUNCH(kernel, grid, block, smem, stream, args);``
  * ``extern __shared__`` -> ``extern`` (one buffer defined in cuda_emu.cpp), ``__shared__`` -> ``static``
    (blocks run one after another);
  * the four PTX helpers of the TMA staging (mbarrier init / expect_tx / try_wait, cp.async.bulk) call the
    emulation's byte-counting phase barrier + memcpy; the lone ``fence.mbarrier_init`` becomes a compiler barrier.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, 'nhd_b200', 'csrc')
GEN = os.path.join(HERE, '_gen')
OUT = os.path.join(HERE, '_emu_cuda.so')
DEPS = [os.path.join(CSRC, f) for f in ('nhd_api.cu', 'nhd_kernels.cuh', 'nhd_core.cuh', 'nhd_ingest.cpp')] + \
    [os.path.join(HERE, f) for f in ('cuda_emu.h', 'cuda_emu.cpp', 'build_emu_cuda.py')] + \
    [os.path.join(ROOT, 'include', 'nhd_b200.h')]

_LAUNCH = re.compile(r'(\b\w+(?:<[^<>;]*>)?)\s*<<<(.*?)>>>\s*\((.*?)\)\s*;', re.S)
_PTX_HELPERS = ('mbar_init', 'mbar_expect_tx', 'mbar_wait', 'tma_load_1d')


def _rewrite_api(text):
    def sub(m):
        cfg = [c.strip() for c in _split_args(m.group(2))]
        while len(cfg) < 4:
            cfg.append('0')
        args = m.group(3).strip()
        return 'EMU_LAUNCH(%s, %s%s);' % (m.group(1), ', '.join(cfg), (', ' + args) if args else '')
    text, n = _LAUNCH.subn(sub, text)
    assert n >= 10, f'only {n} kernel launches rewritten'
    return text.replace('"../../include/nhd_b200.h"', '"%s"' % os.path.join(ROOT, 'include', 'nhd_b200.h'))


def _split_args(s):
    out, depth, cur = [], 0, ''
    for ch in s:
        if ch in '([{':
            depth += 1
        elif ch in ')]}':
            depth -= 1
        if ch == ',' and depth == 0:
            out.append(cur)
            cur = ''
        else:
            cur += ch
    out.append(cur)
    return out


def _rewrite_kernels(text):
    text = text.replace('extern __shared__', 'extern')
    text = re.sub(r'\b__shared__\b', 'static', text)
    text = text.replace('__noinline__', 'EMU_NOINLINE')
    for h in _PTX_HELPERS:
        text, n = re.subn(r'(__device__ __forceinline__ void )%s\(' % h, r'\1ptx_%s(' % h, text)
        assert n == 1, h
    text, n = re.subn(r'asm volatile\("fence\.mbarrier_init[^;]*;"\s*:::\s*"memory"\);', '__asm__ volatile("" ::: "memory");', text)
    assert n == 1
    shim = '''
/* CPU emulation of the TMA staging helpers (tests/emu/cuda_emu.h) */
namespace nhd {
static inline void mbar_init(uint64_t* bar, uint32_t count) { emu_mbar_init(bar, count); }
static inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { emu_mbar_expect_tx(bar, bytes); }
static inline void mbar_wait(uint64_t* bar, uint32_t parity) { emu_mbar_wait(bar, parity); }
static inline void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) { emu_tma_load_1d(dst, src, bytes, bar); }
}
'''
    return text.replace('#include "nhd_core.cuh"\n', '#include "nhd_core.cuh"\n' + shim, 1)


def build(force=False, verbose=False, asan=False, ubsan=False):
    """``asan``: a second library with AddressSanitizer (out-of-bounds accesses to emulated device memory);
    run it with LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0.
    ``ubsan``: a library that aborts on misaligned accesses (a uint4 load from an address that is not a multiple of 16
    is a fault on the GPU and silent on x86) and on out-of-range indices into fixed-size arrays."""
    OUT = os.path.join(HERE, '_emu_cuda_asan.so' if asan else '_emu_cuda_ubsan.so' if ubsan else '_emu_cuda.so')
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in DEPS):
        return OUT
    os.makedirs(GEN, exist_ok=True)
    inc = os.path.join(ROOT, 'include', 'nhd_b200.h')
    with open(os.path.join(GEN, 'nhd_api.cpp'), 'w') as f:
        f.write(_rewrite_api(open(os.path.join(CSRC, 'nhd_api.cu')).read()))
    with open(os.path.join(GEN, 'nhd_kernels.cuh'), 'w') as f:
        f.write(_rewrite_kernels(open(os.path.join(CSRC, 'nhd_kernels.cuh')).read()))
    with open(os.path.join(GEN, 'nhd_core.cuh'), 'w') as f:
        f.write(open(os.path.join(CSRC, 'nhd_core.cuh')).read().replace('"../../include/nhd_b200.h"', '"%s"' % inc)
                .replace('__noinline__', 'EMU_NOINLINE'))
    with open(os.path.join(GEN, 'nhd_ingest.cpp'), 'w') as f:
        f.write(open(os.path.join(CSRC, 'nhd_ingest.cpp')).read().replace('"../../include/nhd_b200.h"', '"%s"' % inc))
    cmd = ['g++', '-O1', '-g', '-std=c++17', '-fPIC', '-shared', '-ffp-contract=off', '-w'] + (['-DNHD_DEBUG_PRINT'] if os.environ.get('EMU_DEBUG_PRINT') else []) + \
          (['-fsanitize=address', '-fno-omit-frame-pointer'] if asan else []) + \
          (['-fsanitize=alignment,bounds', '-fno-sanitize-recover=all'] if ubsan else []) + [
           '-I', os.path.join(HERE, 'fake_cuda'), '-I', GEN, '-o', OUT,
           os.path.join(GEN, 'nhd_api.cpp'), os.path.join(GEN, 'nhd_ingest.cpp'), os.path.join(HERE, 'cuda_emu.cpp'), '-ldl']
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0 or verbose:
        sys.stderr.write(res.stdout + res.stderr[-6000:])
    if res.returncode != 0:
        raise RuntimeError('g++ failed building the emulated library')
    return OUT


if __name__ == '__main__':
    print(build(force=True, verbose=True, asan='--asan' in sys.argv, ubsan='--ubsan' in sys.argv))
