"""Stand-in for the third-party `colorlog` package, which the reference imports
(nhd/NHDCommon.py:2,31) but which is not installed here.  Test infrastructure
only: lets the UNMODIFIED reference modules import so they can serve as the
live oracle / golden-vector generator."""
import logging


class ColoredFormatter(logging.Formatter):
    def __init__(self, fmt=None, datefmt=None, style='%', log_colors=None, **kw):
        super().__init__(fmt.replace('%(log_color)s', '') if fmt else fmt, datefmt)
