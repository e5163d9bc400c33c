// cli_params.h — `SOAPnuke filter` / `SOAPnuke filtersRNA` command line and config-file surface (process_argv.cpp:72-917,
// 1158-1638, defaults global_parameter.h:20-83), parsed once into an immutable HostParams and the
// engine's snk_params POD.
#ifndef SNK_CLI_PARAMS_H
#define SNK_CLI_PARAMS_H
#include <string>
#include <vector>
#include "../../include/snk_engine.h"

namespace snk {

struct HostParams {
    std::string module_name;
    std::string fq1_path, fq2_path, clean_fq1, clean_fq2, output_dir, log = "log";
    std::string trim_fq1, trim_fq2;    // config keys trimFq1= / trimFq2=: every record after trimming, gzip only
    std::string seq_type = "0", output_file_type = "fastq";
    bool input_gz = true, clean_gz = true;
    bool ada_trim = false;
    std::vector<std::string> ada1s, ada2s;
    std::string adapter2_seq;
    int quality_phred = 33, out_quality_phred = 33, low_qual = 5;
    float low_qual_ratio = 0.5f;
    int mean_quality = -1;
    std::string trim_bad_head, trim_bad_tail, trim;
    int max_base_quality = 42;
    float n_ratio = 0.05f, highA_ratio = -1.0f, polyG_tail = -1.0f;
    int polyX_num = -1;
    bool pe_info = false, index_remove = false;
    int threads_requested = 6, threads = 6, patch_size = 0;
    int max_read_length = -1, min_read_length = 30;
    int ada_mis = 2, ada_mis2 = 2, ada_edge = 6, ada_edge2 = 6;
    float ada_mr = 0.5f, ada_mr2 = 0.5f;
    bool is_pe = false;
    std::string contam1, contam2, ct_match_r = "0.2";   // config keys contam1= / contam2= / ctMatchR= (lists: comma separated)
    bool contam_trim = false;          // config key contam_trim: no discard (the trim itself is commented out in 2.1.9)
    std::string global_contams, g_mrs, g_mms;          // config keys global_contams= / glob_cotm_mR= / glob_cotm_mM=
    std::string tile, fov;             // config keys tile= / fov= (removal lists, comma separated)
    // filtersRNA module (global_parameter.h:54-58)
    bool srna = false;
    int ada_rctg = 6, ada_rma = 5, ada_rmm = 4;
    float ada_rar = 0.8f, ada_rer = 0.4f;
    // engine-side knobs (not part of the reference CLI; environment SNK_GPUS / SNK_BATCH_READS)
    int n_gpus = 1;
    unsigned batch_reads = 1u << 15;
    bool fast_exit = false;            // set by main(): skip freeing device / pinned memory, _exit after the reports are written
};

// Parses argv exactly like global_parameter_initial + check_parameter. Returns 0 = run,
// 1 = help/version printed (exit 0), and calls exit(1) after printing "Error:..." on bad input,
// which is the reference's error convention.
int parse_command_line(int argc, char** argv, HostParams& hp);
// HostParams -> snk_params (adapters, trims, logical-thread partition)
void to_engine_params(const HostParams& hp, snk_params& p);
void print_usage(const std::string& module);
void print_version();

}
#endif
