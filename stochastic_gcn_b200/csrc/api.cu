// libsgcn_b200.so -- library-wide state: error string, launch counter, ABI version.
#include "common.cuh"

namespace sgcn {

static thread_local std::string t_last_error;
std::atomic<int64_t> g_launches{0};
unsigned long long* g_trace = nullptr;
int g_pdl = 1;      // on by default; SGCN_TUNE_PDL / env SGCN_PDL=0 turn it off
thread_local int t_pdl_off = 0;
int g_wb_late_trigger = 1;
int g_hist_l2[2] = {0, 0}, g_stream_l2[2] = {0, 0};   // L2 eviction policies {kind, percent} (sgcn_tune_set)
thread_local ShardMap t_hist_map{}, t_feat_map{};

void set_error(const std::string& msg) { t_last_error = msg; }

}  // namespace sgcn

extern "C" {

int sgcn_abi_version(void) { return 1; }

const char* sgcn_last_error(void) { return sgcn::t_last_error.c_str(); }

int64_t sgcn_launch_count(void) { return sgcn::g_launches.load(std::memory_order_relaxed); }

int sgcn_shard_set(int32_t which, int32_t world, int32_t rows_per_shard, const void* const* bases) {
    SGCN_REQUIRE(which == 0 || which == 1, "shard_set: which is 0 (history) or 1 (features)");
    sgcn::ShardMap& m = which == 0 ? sgcn::t_hist_map : sgcn::t_feat_map;
    if (world <= 1) {
        m = sgcn::ShardMap{};
        return SGCN_OK;
    }
    SGCN_REQUIRE(world <= 16 && rows_per_shard > 0 && bases, "shard_set: 2..16 shards of at least one row");
    m.world = world;
    m.rows = rows_per_shard;
    for (int r = 0; r < world; ++r) {
        SGCN_REQUIRE(bases[r] && (((uintptr_t)bases[r]) & 15) == 0, "shard_set: shard bases must be 16-byte aligned");
        m.base[r] = (const float*)bases[r];
    }
    return SGCN_OK;
}

int sgcn_trace_set(void* buf16) {
    sgcn::g_trace = (unsigned long long*)buf16;
    return SGCN_OK;
}

}  // extern "C"
