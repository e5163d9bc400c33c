// cli_params.cpp — command line + config file of the `filter` module.
// Flag surface, defaults, derived values and error texts follow process_argv.cpp (getopt table
// :77-170, option handling :181-520, derived values :533-544, check_parameter :554-917, config
// file :1158-1638). Options that select parts of the reference this engine does not implement
// (rmdup, output split/subsample, streaming, stLFR, base conversion) are rejected with an explicit
// error instead of being silently ignored.
#include "cli_params.h"
#include <sys/stat.h>
#include <cmath>
#include <getopt.h>
#include <sys/sysinfo.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <set>

namespace snk {

namespace {

[[noreturn]] void die(const std::string& msg)
{
    std::cerr << "Error:" << msg << std::endl;
    exit(1);
}
bool ends_with_gz(const std::string& s) { return s.size() >= 3 && s.compare(s.size() - 3, 3, ".gz") == 0; }
std::vector<std::string> split(const std::string& s, char sep)
{
    std::vector<std::string> out;
    std::string cur;
    for (char ch : s) {
        if (ch != sep) cur += ch;
        else { out.push_back(cur); cur.clear(); }
    }
    out.push_back(cur);
    return out;
}
std::string strip(const std::string& s)
{
    size_t a = 0, b = s.size();
    while (a < b && isspace((unsigned char)s[a])) a++;
    while (b > a && isspace((unsigned char)s[b - 1])) b--;
    return s.substr(a, b - a);
}
bool file_exists_and_not_empty(const std::string& path)
{
    // a FIFO / device can be opened only once without losing data (and opening it blocks until the other end is there):
    // its existence is all that is checked here
    struct stat st;
    if (stat(path.c_str(), &st) == 0 && !S_ISREG(st.st_mode) && !S_ISDIR(st.st_mode)) return true;
    std::ifstream f(path.c_str());
    return f && f.peek() != EOF;
}

// -f / -r: a literal adapter, or a file with one adapter per line (process_argv.cpp:242-304)
void read_adapter_option(const char* arg, std::vector<std::string>& list, std::string* literal, int which)
{
    std::ifstream f(arg);
    if (!f) {
        const std::string ada(arg);
        for (char ch : ada)
            if (!strchr("ACGTacgtNn", ch)) {
                std::cerr << "Error:invalid character found in adapter:" << ch << ". Only ACGTacgtNn are supported" << std::endl;
                exit(1);
            }
        list.push_back(ada);
        if (literal) *literal = ada;
        return;
    }
    std::cout << "input adapter" << which << " list file:" << arg << std::endl;
    std::string line;
    while (std::getline(f, line)) list.push_back(line);
}

int phred_code(const std::string& v)
{
    int q = atoi(v.c_str());
    if (q == 1) return 64;
    if (q == 2) return 33;
    return q;
}

void unsupported(const std::string& what) { die("parameter " + what + " selects a part of SOAPnuke this GPU filter engine does not implement"); }

// process_argv.cpp:1158-1638
void init_from_config(HostParams& hp, const char* path)
{
    static const std::set<std::string> bools = {"index", "pe_info", "contam_trim", "notCutNoLFR", "inputAsList", "tenX", "rmdup"};
    static const std::set<std::string> legal = {
        "trimFq1", "trimFq2", "seqType", "outFileType", "contam_trim", "contam1", "contam2", "ctMatchR", "global_contams",
        "glob_cotm_mR", "glob_cotm_mM", "tile", "fov", "index", "qualSys", "outQualSys", "baseConvert", "maxBaseQuality", "overlap",
        "mis", "pe_info", "patch", "maxReadLen", "adaMis", "adaMR", "adaEdge", "adaRCtg", "adaRAr", "adaRMa", "adaREr", "adaRMm",
        "log", "totalReadsNum", "cleanOutSplit", "trim", "trimBadHead", "trimBadTail", "barcodeListPath", "barcodeRegionStr",
        "notCutNoLFR", "inputAsList", "tenX", "rmdup"};
    std::ifstream f(path);
    if (!f) die(std::string("cannot open such file,") + path);
    std::string line;
    while (std::getline(f, line)) {
        if (line.find("#") == 0) continue;
        std::string key, value;
        if (line.find("=") != std::string::npos) {
            std::vector<std::string> e = split(line, '=');
            if (e.size() != 2) die("unrecgonized format parameter," + line);
            key = strip(e[0]); value = strip(e[1]);
        } else {
            key = line;
            if (!bools.count(key)) die("this parameter should set a value," + key);
        }
        if (!legal.count(key)) die("no such parameter," + key);
        auto pair_int = [&](int& a, int& b) {
            if (value.find(",") == std::string::npos) { a = atoi(value.c_str()); b = a; }
            else { auto v = split(value, ','); if (v.size() < 2) die("expected two values in -M parameter"); a = atoi(v[0].c_str()); b = atoi(v[1].c_str()); }
        };
        if (key == "seqType") hp.seq_type = value;
        else if (key == "outFileType") hp.output_file_type = value;
        else if (key == "index") hp.index_remove = true;
        else if (key == "qualSys") hp.quality_phred = phred_code(value);
        else if (key == "outQualSys") hp.out_quality_phred = phred_code(value);
        else if (key == "maxBaseQuality") hp.max_base_quality = atoi(value.c_str());
        else if (key == "pe_info") hp.pe_info = true;
        else if (key == "patch") hp.patch_size = atoi(value.c_str());
        else if (key == "maxReadLen") hp.max_read_length = atoi(value.c_str());
        else if (hp.srna && (key == "adaMis" || key == "adaMR" || key == "adaEdge"))       // process_argv.cpp:1380,1403,1426 + :774-782
            die(std::string("these parameters should not appear in the module,") +
                (key == "adaMis" ? "-M|--adaMis" : key == "adaMR" ? "-A|adaMR" : "-9|--adaEdge"));
        else if (key == "adaMis") pair_int(hp.ada_mis, hp.ada_mis2);
        else if (key == "adaEdge") pair_int(hp.ada_edge, hp.ada_edge2);
        else if (key == "adaMR") {
            if (value.find(",") == std::string::npos) { hp.ada_mr = (float)atof(value.c_str()); hp.ada_mr2 = hp.ada_mr; }
            else { auto v = split(value, ','); if (v.size() < 2) die("expected two values in -A parameter"); hp.ada_mr = (float)atof(v[0].c_str()); hp.ada_mr2 = (float)atof(v[1].c_str()); }
        }
        else if (key == "trimFq1") hp.trim_fq1 = value;                                    // process_argv.cpp:1262-1277
        else if (key == "trimFq2") hp.trim_fq2 = value;
        else if (key == "contam1") hp.contam1 = value;                                     // process_argv.cpp:1286-1300
        else if (key == "contam2") hp.contam2 = value;
        else if (key == "ctMatchR") hp.ct_match_r = value;
        else if (key == "contam_trim") hp.contam_trim = true;
        else if (key == "global_contams") hp.global_contams = value;                       // process_argv.cpp:1302-1312
        else if (key == "glob_cotm_mR") hp.g_mrs = value;
        else if (key == "glob_cotm_mM") hp.g_mms = value;
        else if (key == "tile") hp.tile = value;                                           // process_argv.cpp:1314-1320
        else if (key == "fov") hp.fov = value;
        else if (key == "log") hp.log = value;
        else if (key == "trim") hp.trim = value;
        else if (key == "trimBadHead") hp.trim_bad_head = value;
        else if (key == "trimBadTail") hp.trim_bad_tail = value;
        else if (key == "adaRCtg" || key == "adaRAr" || key == "adaRMa" || key == "adaREr" || key == "adaRMm") {
            if (!hp.srna) {                                       // filtersRNA-only (process_argv.cpp:1447-1470 + :763-770)
                const char* flag = key == "adaRCtg" ? "-S|--" : key == "adaRAr" ? "-s|--" : key == "adaRMa" ? "-U|--" : key == "adaREr" ? "-u|--" : "-b|--";
                die(std::string("these parameters should not appear in the module,") + flag + key);
            }
            if (key == "adaRCtg") hp.ada_rctg = atoi(value.c_str());                               // process_argv.cpp:1447-1470
            else if (key == "adaRAr") hp.ada_rar = (float)atof(value.c_str());
            else if (key == "adaRMa") hp.ada_rma = atoi(value.c_str());
            else if (key == "adaREr") hp.ada_rer = (float)atof(value.c_str());
            else hp.ada_rmm = atoi(value.c_str());
        }
        else unsupported(key);
    }
}

} // namespace

void print_version() { std::cout << "SOAPnuke filter (b200 engine), reference-compatible with 2.1.9" << std::endl; }

void print_usage(const std::string& module)
{
    std::cout << "Usage: SOAPnuke " << module << " [OPTION]...\n"
              << "  -1, --fq1 FILE          fastq1 (.gz or plain)          -2, --fq2 FILE        fastq2 (PE)\n"
              << "  -C, --cleanFq1 NAME     clean fastq1 name in outDir    -D, --cleanFq2 NAME   clean fastq2 name\n"
              << "  -o, --outDir DIR        output directory               -c, --configFile FILE uncommon parameters (key=value)\n"
              << "  -f, --adapter1 SEQ|FILE adapter of fq1 (or list file)  -r, --adapter2 SEQ|FILE adapter of fq2\n"
              << "  -J, --ada_trim          trim adapters instead of discarding the read\n"
              << "  -l, --lowQual INT [5]   -q, --qualRate FLOAT [0.5]   -m, --mean INT [-1]   -n, --nRate FLOAT [0.05]\n"
              << "  -p, --highA FLOAT [-1]  -g, --polyG_tail INT [-1]    -X, --polyX INT [-1]  -4, --minReadLen INT [30]\n"
              << "  -x, --trimBadHead THR,MAXLEN   -y, --trimBadTail THR,MAXLEN   -t, --trim H1,T1[,H2,T2]\n"
              << "  -T, --thread INT [6]    logical worker partition of the reference (statistics/ordering parity) and host threads\n"
              << "  -h, --help   -v, --version\n"
              << "config file keys: seqType outFileType index qualSys outQualSys maxBaseQuality pe_info patch maxReadLen\n"
              << "                  adaMis adaMR adaEdge trim trimBadHead trimBadTail log tile fov\n"
              << "                  contam1 contam2 ctMatchR contam_trim global_contams glob_cotm_mR glob_cotm_mM trimFq1 trimFq2\n"
              << "filtersRNA: -f 5' adapter, -r 3' adapter, defaults minReadLen 18 / maxReadLen 49; config keys adaRCtg adaRAr adaRMa adaREr adaRMm\n"
              << "environment: SNK_GPUS=<n> (GPUs to shard batches over), SNK_BATCH_READS=<n>,\n"
              << "             SNK_KEEP_DEFERRED=1 (write the last deferred batch of a plain-text PE run that the reference\n"
              << "             loses in its final concat; default: drop it like the reference, with a warning on stderr),\n"
              << "             SNK_GZ_CODEC=zlib (.gz outputs through zlib level 2 instead of the in-tree fast deflate encoder),\n"
              << "             SNK_GZ_SERIAL=1 (.gz inputs through one gzread stream instead of the parallel member reader),\n"
              << "             SNK_READ_THREADS / SNK_WRITE_THREADS (page-cache copy threads per input / for the outputs),\n"
              << "             SNK_PREFETCH_MB=<n> (plain inputs: MiB per file read ahead while the CUDA contexts are created; default 2048, 0 = off)\n";
}

int parse_command_line(int argc, char** argv, HostParams& hp)
{
    static const char* short_opts = "j1:2:C:D:o:c:E:Jf:r:l:q:m:x:y:n:p:g:X:t:T:3:4:L:w:hv";
    static const struct option long_opts[] = {
        {"streaming", 0, nullptr, 'j'}, {"fq1", 1, nullptr, '1'}, {"fq2", 1, nullptr, '2'}, {"cleanFq1", 1, nullptr, 'C'},
        {"cleanFq2", 1, nullptr, 'D'}, {"outDir", 1, nullptr, 'o'}, {"configFile", 1, nullptr, 'c'}, {"ref", 1, nullptr, 'E'},
        {"ada_trim", 0, nullptr, 'J'}, {"adapter1", 1, nullptr, 'f'}, {"adapter2", 1, nullptr, 'r'}, {"lowQual", 1, nullptr, 'l'},
        {"qualRate", 1, nullptr, 'q'}, {"mean", 1, nullptr, 'm'}, {"trimBadHead", 1, nullptr, 'x'}, {"trimBadTail", 1, nullptr, 'y'},
        {"nRate", 1, nullptr, 'n'}, {"highA", 1, nullptr, 'p'}, {"polyG_tail", 1, nullptr, 'g'}, {"polyX", 1, nullptr, 'X'},
        {"trim", 1, nullptr, 't'}, {"thread", 1, nullptr, 'T'}, {"minReadLen", 1, nullptr, '4'}, {"output_clean", 1, nullptr, 'w'},
        {"help", 0, nullptr, 'h'}, {"version", 0, nullptr, 'v'}, {nullptr, 0, nullptr, 0}};
    hp.module_name = argv[1];
    if (hp.module_name == "filtersRNA") {        // process_argv.cpp:174-178
        hp.srna = true;
        hp.min_read_length = 18;
        hp.max_read_length = 49;
    }
    int opt;
    while ((opt = getopt_long(argc, argv, short_opts, long_opts, nullptr)) != -1) {
        switch (opt) {
            case '1': hp.fq1_path = optarg; hp.input_gz = ends_with_gz(hp.fq1_path); break;
            case '2': hp.fq2_path = optarg; break;
            case 'C': hp.clean_fq1 = optarg; hp.clean_gz = ends_with_gz(hp.clean_fq1); break;
            case 'D': hp.clean_fq2 = optarg; break;
            case 'o': hp.output_dir = optarg; break;
            case 'J': hp.ada_trim = true; break;
            case 'f': read_adapter_option(optarg, hp.ada1s, nullptr, 1); break;
            case 'r': read_adapter_option(optarg, hp.ada2s, &hp.adapter2_seq, 2); break;
            case 'c': init_from_config(hp, optarg); break;
            case 'l': hp.low_qual = atoi(optarg); break;
            case 'q': hp.low_qual_ratio = (float)atof(optarg); break;
            case 'm': hp.mean_quality = atoi(optarg); break;
            case 'x': hp.trim_bad_head = optarg; break;
            case 'y': hp.trim_bad_tail = optarg; break;
            case 'n': hp.n_ratio = (float)atof(optarg); break;
            case 'p': hp.highA_ratio = (float)atof(optarg); break;
            case 'g': hp.polyG_tail = (float)atof(optarg); break;
            case 'X': hp.polyX_num = (int)atof(optarg); break;
            case 't': hp.trim = optarg; break;
            case 'T': hp.threads_requested = atoi(optarg); break;
            case '4': hp.min_read_length = atoi(optarg); break;
            case 'j': unsupported("-j|--streaming"); break;
            case 'E': unsupported("-E|--ref"); break;
            case 'w': unsupported("-w|--output_clean"); break;
            case '3': case 'L': exit(1);
            case 'v': print_version(); return 1;
            case 'h': print_usage(hp.module_name); return 1;
            default: exit(1);
        }
    }
    if (argc != optind + 1) die("please check the options");
    if (hp.log.find("/") == std::string::npos) hp.log = hp.output_dir + "/" + hp.log;
    if (hp.threads_requested < 1) die("thread number should be a positive integer");
    if (hp.patch_size == 0) hp.patch_size = hp.threads_requested * 20000 / 8;       // process_argv.cpp:541-544

    // ---- check_parameter (process_argv.cpp:554-917)
    if (hp.fq1_path.empty() || !file_exists_and_not_empty(hp.fq1_path)) die("input fastq1 is required");
    if (hp.output_dir.empty()) die("output directory is required");
    if (!hp.fq2_path.empty()) {
        hp.is_pe = true;
        if (!file_exists_and_not_empty(hp.fq2_path)) die("input fastq2 is required");
        if (hp.fq1_path == hp.fq2_path) die("input fq1 and fq2 are the same,please check the parameters");
    }
    if (hp.clean_fq1.empty()) die("output clean fastq is required");
    if (!hp.is_pe && !hp.trim_fq2.empty()) die("input file is not pe data");                      // process_argv.cpp:635
    if (!hp.trim_fq1.empty() || !hp.trim_fq2.empty()) {
        if (hp.trim_fq1.empty() || (hp.is_pe && hp.trim_fq2.empty())) die("trimFq1 and trimFq2 are both required to write the trim files");
        // only the gzip branch works in 2.1.9: the plain-text branch opens its files into the clean-file handles
        // (peprocess.cpp:1784-1789) and peWrite is always given the gzFile handles (:1940)
        if (!ends_with_gz(hp.trim_fq1) || (hp.is_pe && !ends_with_gz(hp.trim_fq2))) die("trim fq file names must end with .gz (gz format)");
    }
    if (hp.is_pe) {
        if (hp.clean_fq2.empty()) die("output clean fastq2 is required");
        if (ends_with_gz(hp.clean_fq1) != ends_with_gz(hp.clean_fq2)) die("the format of clean fastq1 is inconsistent with fastq2");
        if (ends_with_gz(hp.fq1_path) != ends_with_gz(hp.fq2_path)) die("the format of input fastq1 is inconsistent with fastq2");
    } else {
        if (!hp.srna && !hp.adapter2_seq.empty()) die("no need adapter2");       // process_argv.cpp:625
        if (!hp.clean_fq2.empty()) die("input file is not pe data");
    }
    if (hp.srna) {
        if (hp.is_pe) die("filtersRNA with paired input is not served by the GPU filter engine (seProcess path only)");
        // sRNA_hasAdapter starts at adapter offset adptLen-adaRCtg (read_filter.cpp:872): negative = out-of-bounds read in the reference
        if (!hp.ada1s.empty() && (int)hp.ada1s[0].size() < hp.ada_rctg) die("adapter1 is shorter than adaRCtg");
    }
    if (hp.seq_type != "0" && hp.seq_type != "1") die("seq_type value should be 0 or 1");
    if (hp.output_file_type != "fastq" && hp.output_file_type != "fasta") die("output_file_type value should be fastq or fasta");
    if (!hp.tile.empty()) {                     // process_argv.cpp:717-751
        // a '-' sends the reference into `for (size_type ix = size-1; ix >= 0; ix--)` (never terminates, reads out of
        // bounds); ranges never match a tile anyway (check_tile_or_fov compares the tile with the whole "a-b" string)
        if (hp.tile.find("-") != std::string::npos) die("tile ranges (a-b) are not usable in SOAPnuke 2.1.9 (process_argv.cpp:724 never terminates); list the tiles");
        for (char ch : hp.tile)
            if (!isalnum((unsigned char)ch) && ch != ',') die("tile value format error");
    }
    if (!hp.fov.empty()) {
        if (hp.seq_type != "0") { std::cerr << "Warning:Zebra-500 data(--fov), --seqType is 0" << std::endl; exit(1); }   // read_filter.cpp:131-134
        if (hp.fov.find("-") != std::string::npos) die("input tile parameter format error," + hp.fov);
    }
    if (hp.quality_phred != 64 && hp.quality_phred != 33) die("qualityPhred value error");
    if (hp.out_quality_phred != 64 && hp.out_quality_phred != 33) die("outputQualityPhred value error");
    if (!hp.trim.empty()) {
        if (split(hp.trim, ',').size() != (hp.is_pe ? 4u : 2u)) die("trim value format error");
        for (char ch : hp.trim)
            if (!isdigit((unsigned char)ch) && ch != ',') {
                std::cerr << "Error:trim value format error:" << hp.trim << std::endl << "e.g.: -t 10 2 10 2" << std::endl;
                exit(1);
            }
    }
    if (!hp.trim_bad_head.empty() && split(hp.trim_bad_head, ',').size() != (hp.is_pe ? 2u : 1u)) die("trimBadHead value format error");
    if (!hp.trim_bad_tail.empty() && split(hp.trim_bad_tail, ',').size() != (hp.is_pe ? 2u : 1u)) die("trimBadTail value format error");
    if (!hp.is_pe && (!hp.trim_bad_head.empty() || !hp.trim_bad_tail.empty())) {
        // SE accepts only the 1-element form here and then fails in fastq_trim (read_filter.cpp:395-398)
        std::cerr << "Error:low quality base at end format error," << hp.trim_bad_head << " " << hp.trim_bad_head << std::endl;
        exit(1);
    }
    hp.threads = hp.threads_requested;
    if (hp.threads > get_nprocs()) {                       // process_argv.cpp:905-910
        hp.threads = get_nprocs();
        std::cerr << "Warning:threads number exceeds the system cpu number" << std::endl;
    }
    if (hp.patch_size > 5000000) die("patchSize cannot exceed 5M considering memory usage");
    if (hp.patch_size < 1) die("patchSize should be a positive integer");
    if (hp.threads > 160) die("thread number above 160 makes the reference's patch (160/T) zero");
    if (hp.ada1s.size() > SNK_MAX_ADAPTERS || hp.ada2s.size() > SNK_MAX_ADAPTERS) die("too many adapters in the adapter list");
    for (const auto* lst : {&hp.ada1s, &hp.ada2s})
        for (const auto& a : *lst)
            if (a.size() >= SNK_MAX_ADAPTER_LEN) die("adapter longer than the supported maximum");
    if (const char* g = getenv("SNK_GPUS")) hp.n_gpus = atoi(g) > 0 ? atoi(g) : 1;
    if (const char* b = getenv("SNK_BATCH_READS")) hp.batch_reads = atoi(b) > 0 ? (unsigned)atoi(b) : hp.batch_reads;
    return 0;
}

void to_engine_params(const HostParams& hp, snk_params& p)
{
    memset(&p, 0, sizeof(p));
    p.abi_version = SNK_ABI_VERSION;
    p.is_pe = hp.is_pe;
    p.quality_phred = hp.quality_phred; p.out_quality_phred = hp.out_quality_phred;
    p.low_qual = hp.low_qual; p.low_qual_ratio = hp.low_qual_ratio; p.mean_quality = hp.mean_quality;
    p.n_ratio = hp.n_ratio; p.highA_ratio = hp.highA_ratio; p.polyG_tail = hp.polyG_tail; p.polyX_num = hp.polyX_num;
    p.min_read_length = hp.min_read_length; p.max_read_length = hp.max_read_length;
    p.ada_trim = hp.ada_trim;
    p.ada_mis[0] = hp.ada_mis; p.ada_mis[1] = hp.ada_mis2;
    p.ada_mr[0] = hp.ada_mr; p.ada_mr[1] = hp.ada_mr2;
    p.ada_edge[0] = hp.ada_edge; p.ada_edge[1] = hp.ada_edge2;
    const std::vector<std::string>* lists[2] = {&hp.ada1s, &hp.ada2s};
    for (int m = 0; m < 2; m++) {
        p.n_adapters[m] = (int)lists[m]->size();
        for (size_t i = 0; i < lists[m]->size(); i++) {
            p.adapter_len[m][i] = (int)(*lists[m])[i].size();
            memcpy(p.adapter[m][i], (*lists[m])[i].data(), (*lists[m])[i].size());
        }
    }
    if (!hp.trim.empty()) {                     // peprocess.cpp:1692-1699 / get_se_hard_trim
        auto e = split(hp.trim, ',');
        p.has_hard_trim = 1;
        p.hard_head[0] = atoi(e[0].c_str()); p.hard_tail[0] = atoi(e[1].c_str());
        if (e.size() >= 4) { p.hard_head[1] = atoi(e[2].c_str()); p.hard_tail[1] = atoi(e[3].c_str()); }
    }
    if (!hp.trim_bad_head.empty()) {            // read_filter.cpp:393-406
        auto e = split(hp.trim_bad_head, ',');
        p.has_trim_bad_head = 1;
        if (e.size() == 2) { p.bad_head_thr = atoi(e[0].c_str()); p.bad_head_max = atoi(e[1].c_str()); }
    }
    if (!hp.trim_bad_tail.empty()) {
        auto e = split(hp.trim_bad_tail, ',');
        p.has_trim_bad_tail = 1;
        if (e.size() == 2) { p.bad_tail_thr = atoi(e[0].c_str()); p.bad_tail_max = atoi(e[1].c_str()); }
    }
    p.index_remove = hp.index_remove;
    p.max_base_quality = hp.max_base_quality;
    p.contam_discard = hp.contam_trim ? 0 : 1;
    p.contam_trim = hp.contam_trim ? 1 : 0;
    {
        const std::string* lists[2] = {&hp.contam1, &hp.contam2};
        for (int m = 0; m < 2; m++) {
            const std::string& v = *lists[m];
            if (v.empty() || (m == 1 && !hp.is_pe)) continue;
            std::vector<std::string> seqs;
            std::vector<int> thr;
            if (v.find(",") == std::string::npos) {               // hasContam(): double product (read_filter.cpp:609)
                seqs.push_back(v);
                thr.push_back((int)ceil((double)v.size() * atof(hp.ct_match_r.c_str())));
            } else {                                              // hasContams(): one ratio per sequence, float product (:499,:514)
                seqs = split(v, ',');
                if (hp.ct_match_r.find(",") == std::string::npos) die("the number of ctMatchR value should equal to that of contam sequences");
                const std::vector<std::string> mrs = split(hp.ct_match_r, ',');
                if (mrs.size() != seqs.size())
                    die("the number of ctMatchR value should equal to that of contam sequences," + std::to_string(seqs.size()) + ".vs." + std::to_string(mrs.size()));
                for (size_t i = 0; i < seqs.size(); i++) {
                    const float mr = (float)atof(mrs[i].c_str());
                    thr.push_back((int)ceil((int)seqs[i].size() * mr));
                }
            }
            if (seqs.size() > SNK_MAX_CONTAMS) die("too many contaminant sequences");
            p.n_contams[m] = (int)seqs.size();
            for (size_t i = 0; i < seqs.size(); i++) {
                if (seqs[i].size() >= SNK_MAX_ADAPTER_LEN) die("contaminant sequence longer than the supported maximum");
                p.contam_len[m][i] = (int)seqs[i].size();
                p.contam_seg_thr[m][i] = thr[i];
                memcpy(p.contam[m][i], seqs[i].data(), seqs[i].size());
            }
        }
    }
    if (!hp.global_contams.empty()) {            // hasGlobalContams (read_filter.cpp:927-944)
        const std::vector<std::string> seqs = split(hp.global_contams, ','), mrs = split(hp.g_mrs, ','), mms = split(hp.g_mms, ',');
        if (seqs.size() != mrs.size() || seqs.size() != mms.size() || hp.g_mrs.empty() || hp.g_mms.empty())
            die("the number of global contamination sequences should equal to that of related parameters");
        if (seqs.size() > SNK_MAX_CONTAMS) die("too many global contaminant sequences");
        p.n_gcontams = (int)seqs.size();
        for (size_t i = 0; i < seqs.size(); i++) {
            if (seqs[i].empty() || seqs[i].size() >= SNK_MAX_ADAPTER_LEN) die("global contaminant sequence length out of range");
            const float mr = (float)atof(mrs[i].c_str());
            p.gcontam_len[i] = (int)seqs[i].size();
            p.gcontam_min_match[i] = (int)((int)seqs[i].size() * mr);            // int(cl*min_matchRatio) (:970)
            p.gcontam_mismatch[i] = atoi(mms[i].c_str());
            memcpy(p.gcontam[i], seqs[i].data(), seqs[i].size());
        }
    }
    p.seq_type1 = hp.seq_type != "0";
    {
        struct { const std::string* src; int32_t* n; char (*dst)[SNK_ID_FILTER_LEN]; } lists[2] = {{&hp.tile, &p.n_tile, p.tile}, {&hp.fov, &p.n_fov, p.fov}};
        for (auto& l : lists) {
            *l.n = 0;
            if (l.src->empty()) continue;
            for (const std::string& e : split(*l.src, ',')) {
                if (e.empty() || e.size() > SNK_ID_FILTER_LEN) continue;       // can never equal a 4-digit tile / 8-character fov
                if (*l.n >= SNK_MAX_ID_FILTERS) die("too many entries in the tile / fov list");
                memcpy(l.dst[*l.n], e.data(), e.size());
                (*l.n)++;
            }
        }
    }
    p.srna = hp.srna;
    p.ada_rctg = hp.ada_rctg; p.ada_rar = hp.ada_rar; p.ada_rma = hp.ada_rma; p.ada_rer = hp.ada_rer; p.ada_rmm = hp.ada_rmm;
    p.n_slots = hp.threads;                                       // logical reference threads
    p.slot_block = (int64_t)hp.patch_size * (160 / hp.threads);   // peprocess.cpp:81, :2063
}

} // namespace snk
