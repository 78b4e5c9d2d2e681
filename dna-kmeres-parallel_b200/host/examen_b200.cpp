// examen_b200 — command-line driver with the reference's main() flow
// (main.cu:120-174: enumerate k-mers, importSeqs, count, distances, CSV), with the
// hard-coded paths of main.cu:48,178,216 turned into arguments and k a runtime flag.
// "Next" row f3 of SURVEY.md §8; the GPU work goes through the C ABI.
//
//   examen_b200 <input.fasta> [-k K] [--nonl] [--max-seqs N] [--out parallel_results.csv]
//               [--sums sums.txt] [--gpu-parse]
// --gpu-parse: the FASTA file is copied to the device as it is and parsed there (no record limit)
#include <chrono>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <memory>

#include "kmer_b200.hpp"

int main(int argc, char** argv) {
    std::string file, out = "parallel_results.csv", sums_path;
    int k = 3;
    bool nonl = false, gpu_parse = false;
    long max_seqs = 100;  // MAX_SEQS, main.cu:30
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "-k") && i + 1 < argc) k = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--nonl")) nonl = true;
        else if (!strcmp(argv[i], "--gpu-parse")) gpu_parse = true;
        else if (!strcmp(argv[i], "--max-seqs") && i + 1 < argc) max_seqs = atol(argv[++i]);
        else if (!strcmp(argv[i], "--out") && i + 1 < argc) out = argv[++i];
        else if (!strcmp(argv[i], "--sums") && i + 1 < argc) sums_path = argv[++i];
        else if (argv[i][0] != '-') file = argv[i];
        else { fprintf(stderr, "unknown argument %s\n", argv[i]); return 2; }
    }
    if (file.empty()) {
        fprintf(stderr, "usage: %s <input.fasta> [-k K] [--nonl] [--max-seqs N] [--out csv] [--sums txt] [--gpu-parse]\n", argv[0]);
        return 2;
    }
    try {
        std::cout << "K = " << k << std::endl;  // main.cu:145
        std::unique_ptr<kmerb200::Engine> engp;
        if (gpu_parse) engp.reset(new kmerb200::Engine(0));  // otherwise the file is read first, as in the reference
        kmerb200::Sequences seqs = gpu_parse ? kmerb200::importSeqsGpu(*engp, file, nonl)
                                             : nonl ? kmerb200::importSeqsNoNL(file, max_seqs) : kmerb200::importSeqs(file, max_seqs);
        std::cout << "Size all seqs:" << seqs.size_all_seqs << std::endl;          // main.cu:166
        std::cout << seqs.numberOfSequenses << " sequences read ." << std::endl;  // main.cu:167
        printf("\n\aParallel:\n");                                                 // main.cu:170
        if (!engp) engp.reset(new kmerb200::Engine(0));
        kmerb200::Engine& eng = *engp;
        auto t0 = std::chrono::steady_clock::now();
        int32_t* d_sums = nullptr;
        std::vector<int32_t> sums = eng.sumKmereCoincidences(seqs, k, &d_sums);
        auto t1 = std::chrono::steady_clock::now();
        double ms1 = std::chrono::duration<double, std::milli>(t1 - t0).count();
        std::cout << "Elapsed parallel timer step 1: " << ms1 << " ms, " << ms1 / 1000 << " secs" << std::endl;  // main.cu:300
        if (!sums_path.empty()) kmerb200::check(kc_dump_counts(sums_path == "-" ? nullptr : sums_path.c_str(), sums.data(), k, seqs.numberOfSequenses));
        std::vector<float> mins = eng.minKmeres(d_sums, seqs, k);
        kc_device_free(eng.ctx(), d_sums);
        auto t2 = std::chrono::steady_clock::now();
        double ms2 = std::chrono::duration<double, std::milli>(t2 - t1).count();
        std::cout << "Elapsed parallel step 2 timer: " << ms2 << " ms, " << ms2 / 1000 << " secs" << std::endl;      // main.cu:344
        std::cout << "Total time elapsed parallel: " << ms1 + ms2 << " ms, " << (ms1 + ms2) / 1000 << " secs" << std::endl;  // main.cu:350
        kmerb200::check(kc_dump_distances(out.c_str(), mins.data(), mins.size()));  // main.cu:355-358
    } catch (const kmerb200::Error& e) {
        fprintf(stderr, "examen_b200: error %d: %s\n", e.code, e.what());
        return 1;
    }
    return 0;
}
