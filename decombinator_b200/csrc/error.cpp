// error.cpp -- thread-local last-error text and small ABI helpers.
#include "dcb_internal.h"

#include <cstdarg>
#include <cstdio>

static thread_local char g_err[512] = "";

void dcb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" {

const char* dcb_last_error(void) { return g_err; }
int dcb_abi_version(void) { return DCB_ABI_VERSION; }

const char* dcb_counter_name(int i) {
    static const char* names[DCB_NCOUNTERS] = {
        "verr1", "verr2", "jerr1", "jerr2",
        "dcrfilter_intertagN", "dcrfilter_toolong_intertag", "dcrfilter_imposs_deletion", "dcrfilter_tag_overlap",
        "multiple_v_matches", "v_del_failed_tag_at_end", "v_del_failed", "foundv1notv2", "foundv2notv1",
        "no_vtags_found", "multiple_j_matches", "j_del_failed", "foundj1notj2", "foundj2notj1",
        "no_j_assigned", "VJ_assignment_failed",
    };
    return (i >= 0 && i < DCB_NCOUNTERS) ? names[i] : nullptr;
}

}  // extern "C"
