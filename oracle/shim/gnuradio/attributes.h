/* Test infrastructure only (oracle/): minimal stand-in for <gnuradio/attributes.h>
 * so that the UNMODIFIED reference sources under /root/reference/lib compile
 * without GNU Radio. Written from scratch for this repo; not GNU Radio code. */
#ifndef ORACLE_SHIM_GNURADIO_ATTRIBUTES_H
#define ORACLE_SHIM_GNURADIO_ATTRIBUTES_H
#define __GR_ATTR_EXPORT __attribute__((visibility("default")))
#define __GR_ATTR_IMPORT __attribute__((visibility("default")))
#endif
