/* TEST INFRASTRUCTURE (oracle/_ref build only).
 * Stand-in for the cmake-generated iqtree_config.h (reference iqtree_config.h.in)
 * so that /root/reference/tools.h can be included without running the
 * reference's build system. */
#ifndef MPB200_SHIM_IQTREE_CONFIG_H
#define MPB200_SHIM_IQTREE_CONFIG_H
#define iqtree_VERSION_MAJOR 1
#define iqtree_VERSION_MINOR 1
#define iqtree_VERSION_PATCH 0
#define HAVE_GETTIMEOFDAY
#define HAVE_GETRUSAGE
#endif
