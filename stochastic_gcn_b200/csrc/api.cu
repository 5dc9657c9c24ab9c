// libsgcn_b200.so -- library-wide state: error string, launch counter, ABI version.
#include "common.cuh"

namespace sgcn {

static thread_local std::string t_last_error;
std::atomic<int64_t> g_launches{0};
unsigned long long* g_trace = nullptr;
int g_pdl = 1;      // on by default; SGCN_TUNE_PDL / env SGCN_PDL=0 turn it off
thread_local int t_pdl_off = 0;

void set_error(const std::string& msg) { t_last_error = msg; }

}  // namespace sgcn

extern "C" {

int sgcn_abi_version(void) { return 1; }

const char* sgcn_last_error(void) { return sgcn::t_last_error.c_str(); }

int64_t sgcn_launch_count(void) { return sgcn::g_launches.load(std::memory_order_relaxed); }

int sgcn_trace_set(void* buf16) {
    sgcn::g_trace = (unsigned long long*)buf16;
    return SGCN_OK;
}

}  // extern "C"
