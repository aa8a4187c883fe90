/* hand-written stand-in for the configure-generated libavutil/avconfig.h
 * (oracle/_ref build only; see ../config.h) */
#ifndef AVUTIL_AVCONFIG_H
#define AVUTIL_AVCONFIG_H
#define AV_HAVE_BIGENDIAN 0
#endif
