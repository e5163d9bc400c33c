// host_common.h — small shared helpers of the host side (error slot behind snk_last_error()).
#ifndef SNK_HOST_COMMON_H
#define SNK_HOST_COMMON_H
#include <string>
#include "../../include/snk_engine.h"

namespace snk {
void set_error(const std::string& msg);
const char* last_error();
int report_write_pe(const snk_params& p, const uint64_t* stats, const std::string& dir);
int report_write_se(const snk_params& p, const uint64_t* stats, const std::string& dir);
int params_check(const snk_params& p);
}
#endif
