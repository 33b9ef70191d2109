// dcb_internal.h -- definitions shared by the translation units of libdcb.so (not installed).
#ifndef DCB_INTERNAL_H
#define DCB_INTERNAL_H

#include "../../include/dcb.h"
#include "dcb_tables.h"

#include <vector>

struct dcb_tagset {
    int n_tags = 0, split = 0, is_v = 0;
    std::vector<int> tag_len;
    std::vector<uint32_t> general;  // blob for the general (fallback) kernel
    std::vector<uint32_t> fast;     // blob for the exact-tag kernel
};

#if defined(__GNUC__)
__attribute__((format(printf, 1, 2)))
#endif
void dcb_set_error(const char* fmt, ...);

#endif
