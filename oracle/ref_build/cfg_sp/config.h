/* stub for the autoconf-generated config.h: single-precision build */
