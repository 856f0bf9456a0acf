/*
 * wmb_framer.h -- internal header of the host-side framers; the exported interface is
 * include/wmbus_b200_framer.h.
 */
#ifndef WMB_FRAMER_H
#define WMB_FRAMER_H

#include "wmbus_b200_framer.h"

#endif
