// ORACLE — TEST INFRASTRUCTURE ONLY.  Command-line front end of the CPU restatement:
//   mbl_oracle classify [--seq-mode 1|2|3] [--threads N] <fastx> [<fastx2>] <dbdir> <outdir> <jobid>
// mirrors `metabuli classify` (src/workflow/classify.cpp:39-200) for the flags the hot path reads.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "mbl_oracle.hpp"

int main(int argc, char **argv) {
    if (argc < 2 || strcmp(argv[1], "classify") != 0) {
        fprintf(stderr, "usage: %s classify [--seq-mode M] [--threads N] [--min-score F] [--min-sp-score F] "
                        "[--tie-ratio F] [--min-cons-cnt N] [--min-cons-cnt-euk N] <fastx> [<fastx2>] <db> <out> <job>\n", argv[0]);
        return 2;
    }
    orc::Options opt;
    std::vector<std::string> pos;
    for (int i = 2; i < argc; ++i) {
        std::string a = argv[i];
        auto need = [&](const char *) { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "--seq-mode") opt.seqMode = atoi(need(""));
        else if (a == "--threads") opt.threads = atoi(need(""));
        else if (a == "--min-score") opt.minScore = (float)atof(need(""));
        else if (a == "--min-sp-score") opt.minSpScore = (float)atof(need(""));
        else if (a == "--tie-ratio") opt.tieRatio = (float)atof(need(""));
        else if (a == "--min-cons-cnt") opt.minConsCnt = atoi(need(""));
        else if (a == "--min-cons-cnt-euk") opt.minConsCntEuk = atoi(need(""));
        else if (a == "--accession-level") opt.accessionLevel = atoi(need(""));
        else if (a == "--max-ram") (void)need("");
        else pos.push_back(a);
    }
    size_t want = opt.seqMode == 2 ? 5 : 4;
    if (pos.size() != want) { fprintf(stderr, "expected %zu positional arguments, got %zu\n", want, pos.size()); return 2; }
    std::string q1 = pos[0], q2 = opt.seqMode == 2 ? pos[1] : "", db = pos[want - 3], out = pos[want - 2], job = pos[want - 1];
    std::string tsv, err;
    size_t nk = 0, nm = 0;
    auto t0 = std::chrono::steady_clock::now();
    if (!orc::classify_files(q1, q2, db, opt, tsv, &err, &nk, &nm)) { fprintf(stderr, "error: %s\n", err.c_str()); return 1; }
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::ofstream f(out + "/" + job + "_classifications.tsv", std::ios::binary);
    f << tsv;
    printf("Query k-mer number     : %zu\nK-mer match count      : %zu\nwall %.3f s\n", nk, nm, sec);
    return 0;
}
