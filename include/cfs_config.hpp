#ifndef CFS_CONFIG_HPP
#define CFS_CONFIG_HPP
// Build-time switches (_USE_DOUBLE, _LOG_INFO); see include/config.h.
#include <config.h>
#endif
