/* config.h -- stands in for the file the reference's configure script
 * generates (configure.ac:58-67). The build switches keep their meaning:
 *   --enable-dp   ->  compile with -DCFS_ENABLE_DP   ->  _USE_DOUBLE
 *   --enable-log  ->  compile with -DCFS_ENABLE_LOG  ->  _LOG_INFO
 * Only consumers read _USE_DOUBLE (bench VALUE typedef); the library always
 * carries both precisions. */
#ifndef CFS_B200_CONFIG_H
#define CFS_B200_CONFIG_H
#if defined(CFS_ENABLE_DP) && !defined(_USE_DOUBLE)
#define _USE_DOUBLE 1
#endif
#if defined(CFS_ENABLE_LOG) && !defined(_LOG_INFO)
#define _LOG_INFO 1
#endif
#define CFS_B200 1
#endif
