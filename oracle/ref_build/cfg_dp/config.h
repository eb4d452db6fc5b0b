/* stub for the autoconf-generated config.h: --enable-dp build */
#define _USE_DOUBLE 1
