/*
 * pyset_model.c — ORACLE (test infrastructure only).  See pyset_model.h.
 *
 * Follows CPython 3.12 Objects/setobject.c and Objects/tupleobject.c; the
 * behaviour is unchanged since 3.8 for the operations modelled here.
 */
#include "pyset_model.h"

#include <string.h>

#define LINEAR_PROBES 9
#define PERTURB_SHIFT 5

uint64_t py_hash_int(int64_t v)
{
    /* long_hash: value mod (2^61 - 1); identity for the small ints used here.
     * hash(-1) == -2 is irrelevant (no negative keys on this path). */
    return (uint64_t)v;
}

uint64_t py_hash_tuple(const int* items, int len)
{
    /* tuplehash(), xxHash-style (CPython >= 3.8) */
    const uint64_t P1 = 11400714785074694791ULL;
    const uint64_t P2 = 14029467366897019727ULL;
    const uint64_t P5 = 2870177450012600261ULL;
    uint64_t acc = P5;
    for (int i = 0; i < len; i++) {
        uint64_t lane = py_hash_int(items[i]);
        acc += lane * P2;
        acc = (acc << 31) | (acc >> 33);
        acc *= P1;
    }
    acc += (uint64_t)len ^ (P5 ^ 3527539ULL);
    if (acc == (uint64_t)-1)
        return 1546275796ULL;
    return acc;
}

uint64_t pyset_tuple_key(const int* items, int len)
{
    uint64_t k = (uint64_t)len << 56;
    for (int i = 0; i < len; i++)
        k |= (uint64_t)(items[i] & 0xF) << (4 * i);
    return k;
}

void pyset_key_tuple(uint64_t key, int* items_out, int* len_out)
{
    int len = (int)(key >> 56);
    for (int i = 0; i < len; i++)
        items_out[i] = (int)((key >> (4 * i)) & 0xF);
    *len_out = len;
}

void pyset_init(pyset* s)
{
    memset(s->slots, 0, sizeof(pyset_entry) * PYSET_MINSIZE);
    s->mask = PYSET_MINSIZE - 1;
    s->fill = 0;
    s->used = 0;
}

/* set_insert_clean(): insert into a table known to have no equal key and no dummies */
static void insert_clean(pyset_entry* table, size_t mask, uint64_t key, uint64_t hash)
{
    size_t perturb = hash;
    size_t i = (size_t)hash & mask;
    for (;;) {
        pyset_entry* e = &table[i];
        if (!e->live)
            goto found_null;
        if (i + LINEAR_PROBES <= mask) {
            for (size_t j = 0; j < LINEAR_PROBES; j++) {
                e++;
                if (!e->live)
                    goto found_null;
            }
        }
        perturb >>= PERTURB_SHIFT;
        i = (i * 5 + 1 + perturb) & mask;
        continue;
found_null:
        e->key = key;
        e->hash = hash;
        e->live = 1;
        return;
    }
}

/* set_table_resize(): smallest power of two > minused, re-insert in old slot order */
static void table_resize(pyset* s, size_t minused)
{
    static _Thread_local pyset_entry old[PYSET_MAX_SLOTS];
    size_t oldmask = s->mask;
    size_t newsize = PYSET_MINSIZE;
    while (newsize <= minused)
        newsize <<= 1;
    memcpy(old, s->slots, sizeof(pyset_entry) * (oldmask + 1));
    memset(s->slots, 0, sizeof(pyset_entry) * newsize);
    s->mask = newsize - 1;
    for (size_t i = 0; i <= oldmask; i++)
        if (old[i].live)
            insert_clean(s->slots, s->mask, old[i].key, old[i].hash);
    s->fill = s->used;
}

int pyset_add(pyset* s, uint64_t key, uint64_t hash)
{
    /* set_add_entry(); there are never dummy entries here (no deletions) */
    size_t mask = s->mask;
    size_t i = (size_t)hash & mask;
    size_t perturb = hash;
    pyset_entry* e;
    for (;;) {
        e = &s->slots[i];
        int probes = (i + LINEAR_PROBES <= mask) ? LINEAR_PROBES : 0;
        do {
            if (!e->live)
                goto found_unused;
            if (e->hash == hash && e->key == key)
                return 0;                        /* found_active */
            e++;
        } while (probes--);
        perturb >>= PERTURB_SHIFT;
        i = (i * 5 + 1 + perturb) & mask;
    }
found_unused:
    s->fill++;
    s->used++;
    e->key = key;
    e->hash = hash;
    e->live = 1;
    if (s->fill * 5 < mask * 3)
        return 1;
    table_resize(s, s->used > 50000 ? s->used * 2 : s->used * 4);
    return 1;
}

int pyset_contains(const pyset* s, uint64_t key, uint64_t hash)
{
    /* set_lookkey() */
    size_t mask = s->mask;
    size_t i = (size_t)hash & mask;
    size_t perturb = hash;
    for (;;) {
        const pyset_entry* e = &s->slots[i];
        int probes = (i + LINEAR_PROBES <= mask) ? LINEAR_PROBES : 0;
        do {
            if (!e->live)
                return 0;
            if (e->hash == hash && e->key == key)
                return 1;
            e++;
        } while (probes--);
        perturb >>= PERTURB_SHIFT;
        i = (i * 5 + 1 + perturb) & mask;
    }
}

size_t pyset_list(const pyset* s, uint64_t* keys_out, uint64_t* hashes_out)
{
    size_t n = 0;
    for (size_t i = 0; i <= s->mask; i++) {
        if (s->slots[i].live) {
            if (keys_out) keys_out[n] = s->slots[i].key;
            if (hashes_out) hashes_out[n] = s->slots[i].hash;
            n++;
        }
    }
    return n;
}

void pyset_intersection(const pyset* a, const pyset* b, pyset* out)
{
    /* set_intersection(so=a, other=b): iterate `other`, unless it is larger */
    const pyset* so = a;
    const pyset* other = b;
    if (other->used > so->used) {
        const pyset* t = so; so = other; other = t;
    }
    pyset_init(out);
    for (size_t i = 0; i <= other->mask; i++) {
        const pyset_entry* e = &other->slots[i];
        if (e->live && pyset_contains(so, e->key, e->hash))
            pyset_add(out, e->key, e->hash);
    }
}
