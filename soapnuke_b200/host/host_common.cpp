// host_common.cpp — error slot and parameter validation shared by the engine and the CLI.
#include "host_common.h"
#include <cstring>

namespace snk {
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
const char* last_error() { return g_err.c_str(); }

// What the hot path needs from the parameters. The reference would divide by zero / index out of
// range on these (read_filter.cpp:714-715); reject them up front instead.
int params_check(const snk_params& p)
{
    if (p.abi_version != SNK_ABI_VERSION) { set_error("snk_params.abi_version mismatch"); return 1; }
    if (p.n_slots < 1 || p.n_slots > SNK_MAX_SLOTS) { set_error("n_slots out of range"); return 1; }
    if (p.slot_block < 1) { set_error("slot_block must be >= 1"); return 1; }
    if (p.quality_phred != 33 && p.quality_phred != 64) { set_error("qualityPhred value error"); return 1; }
    if (p.out_quality_phred != 33 && p.out_quality_phred != 64) { set_error("outputQualityPhred value error"); return 1; }
    for (int m = 0; m < 2; m++) {
        if (p.n_adapters[m] < 0 || p.n_adapters[m] > SNK_MAX_ADAPTERS) { set_error("too many adapters"); return 1; }
        if (p.n_adapters[m] > 0 && p.ada_mis[m] + 1 == 0) { set_error("adaMis must not be -1"); return 1; }
        if (p.n_adapters[m] > 0 && p.ada_edge[m] < 0) { set_error("adaEdge must be >= 0"); return 1; }
        for (int i = 0; i < p.n_adapters[m]; i++) {
            int L = p.adapter_len[m][i];
            if (L < 0 || L >= SNK_MAX_ADAPTER_LEN) { set_error("adapter too long"); return 1; }
            if ((int)strnlen(p.adapter[m][i], SNK_MAX_ADAPTER_LEN) < L) { set_error("adapter_len exceeds adapter string"); return 1; }
        }
    }
    if (p.srna) {
        if (p.is_pe) { set_error("filtersRNA runs on single-end input (seProcess)"); return 1; }
        // sRNA_hasAdapter starts at adapter offset adptLen - adaRCtg (read_filter.cpp:872)
        if (p.n_adapters[0] > 0 && p.adapter_len[0][0] < p.ada_rctg) { set_error("adapter1 is shorter than adaRCtg"); return 1; }
    }
    for (int m = 0; m < 2; m++) {
        if (p.n_contams[m] < 0 || p.n_contams[m] > SNK_MAX_CONTAMS) { set_error("too many contaminant sequences"); return 1; }
        if (p.n_contams[m] > 0 && p.ada_mis[m] + 1 == 0) { set_error("adaMis must not be -1"); return 1; }
        if (p.n_contams[m] > 0 && p.srna) { set_error("contaminant sequences are not part of filtersRNA"); return 1; }
        for (int i = 0; i < p.n_contams[m]; i++) {
            const int L = p.contam_len[m][i];
            if (L < 0 || L >= SNK_MAX_ADAPTER_LEN) { set_error("contaminant sequence too long"); return 1; }
            if ((int)strnlen(p.contam[m][i], SNK_MAX_ADAPTER_LEN) < L) { set_error("contam_len exceeds the contaminant string"); return 1; }
        }
    }
    if (p.n_gcontams < 0 || p.n_gcontams > SNK_MAX_CONTAMS) { set_error("too many global contaminant sequences"); return 1; }
    if (p.n_gcontams > 0 && p.srna) { set_error("global contaminants are not part of filtersRNA"); return 1; }
    for (int i = 0; i < p.n_gcontams; i++) {
        const int L = p.gcontam_len[i];
        if (L <= 0 || L >= SNK_MAX_ADAPTER_LEN || (int)strnlen(p.gcontam[i], SNK_MAX_ADAPTER_LEN) < L) { set_error("global contaminant sequence length out of range"); return 1; }
        if (p.gcontam_min_match[i] < 0 || p.gcontam_min_match[i] > L) { set_error("global contaminant match ratio must be in [0,1]"); return 1; }
        for (int k = 0; k < L; k++) {
            const char ch = (char)(p.gcontam[i][k] & ~0x20);
            if (ch != 'A' && ch != 'C' && ch != 'G' && ch != 'T' && ch != 'N') { set_error("unrecognized base in a global contaminant sequence"); return 1; }   // reversecomplementary()
        }
    }
    if (p.n_tile < 0 || p.n_tile > SNK_MAX_ID_FILTERS || p.n_fov < 0 || p.n_fov > SNK_MAX_ID_FILTERS) { set_error("too many tile / fov entries"); return 1; }
    if (p.n_fov > 0 && p.seq_type1) { set_error("Zebra-500 data(--fov), --seqType is 0"); return 1; }     // read_filter.cpp:131-134
    if (p.has_hard_trim) for (int m = 0; m < 2; m++)
        if (p.hard_head[m] < 0 || p.hard_tail[m] < 0) { set_error("trim value format error"); return 1; }
    return 0;
}
}

extern "C" {
const char* snk_last_error(void) { return snk::last_error(); }
int snk_abi_version(void) { return SNK_ABI_VERSION; }
size_t snk_stats_slot_words(void) { return SNK_SLOT_WORDS; }
int snk_params_check(const snk_params* p) { if (!p) { snk::set_error("null params"); return 1; } return snk::params_check(*p); }
}
