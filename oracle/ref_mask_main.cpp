// Test infrastructure (container-only): masks reads with the REFERENCE's own objects — NucleotideMatrix, ProbabilityMatrix and
// tantan::maskSequences from the reference's static libraries built by oracle/build_ref.sh — following the ten lines of
// SeqIterator::maskLowComplexityRegions (SeqIterator.cpp:154-175).  Reads one sequence per line on stdin, writes the masked
// sequence per line on stdout.  Used by tests/golden/gen_synth_golden.py to write the per-letter golden of --mask 1.
//   usage: ref_mask <mask-prob>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>

#include "NucleotideMatrix.h"
#include "Parameters.h"
#include "tantan.h"

const char* binary_name = "ref_mask";          // the two hooks the reference's framework library expects from its host binary
DEFAULT_PARAMETER_SINGLETON_INIT

int main(int argc, char** argv) {
    const float mask_prob = argc > 1 ? (float)atof(argv[1]) : 0.9f;
    Parameters& par = Parameters::getInstance();
    par.initMatrices();                        // what parseParameters does before a workflow runs (Parameters.cpp:1662-1690)
    NucleotideMatrix sub(par.scoringMatrixFile.values.nucleotide().c_str(), 1.0, 0.0);
    ProbabilityMatrix prob(sub);
    std::string line, out;
    while (std::getline(std::cin, line)) {
        out.assign(line.size(), '\0');
        for (size_t i = 0; i < line.size(); ++i) out[i] = (char)sub.aa2num[static_cast<int>(line[i])];
        tantan::maskSequences((unsigned char*)&out[0], (unsigned char*)&out[0] + out.size(), 50, prob.probMatrixPointers, 0.005, 0.05, 0.9, 0, 0,
                              mask_prob, prob.hardMaskTable);
        for (size_t i = 0; i < line.size(); ++i) out[i] = ((unsigned char)out[i] == prob.hardMaskTable[0]) ? 'N' : line[i];
        std::cout << out << '\n';
    }
    return 0;
}
