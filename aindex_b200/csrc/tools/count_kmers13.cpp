// count_kmers13 <input_file> <hash_file.pf> <output_file.tf.bin> [threads]
// GPU version of the reference tool (src/count_kmers13.cpp:546-612): same positional arguments
// (the worker threads of count_kmers13.cpp:305-309 are GPUs here: every visible GPU is used, at most `threads`
// of them, rounded down to a power of two; AINDEX_CUDA_DEVICE pins a single device), same 4^13 x uint64 output in
// MPHF order, same statistics block.
#include "tool_common.hpp"

int main(int argc, char **argv) {
    if (argc < 4) {
        fprintf(stderr, "Usage: %s <input_file> <hash_file.pf> <output_file.tf.bin> [threads]\n", argv[0]);
        return 1;
    }
    // threads -> GPUs (count_kmers13.cpp:566: num_threads defaults to every hardware thread = every GPU here)
    int want = argc > 4 ? atoi(argv[4]) : 0;
    aix_multi *mg = nullptr;
    int one = tool_device();
    int rc0 = getenv("AINDEX_CUDA_DEVICE") ? aix_multi_create(1, &one, &mg) : aix_multi_create(0, nullptr, &mg);
    if (rc0 != AIX_OK) {
        fprintf(stderr, "Error: %s\n", aix_multi_last_error(nullptr));
        return 10;
    }
    int n_gpu = aix_multi_size(mg);
    if (want > 0 && want < n_gpu) n_gpu = want;
    while (n_gpu & (n_gpu - 1)) n_gpu &= n_gpu - 1;  // the k-mer ranges must divide 4^13
    if (n_gpu != aix_multi_size(mg)) {
        aix_multi_destroy(mg);
        if (aix_multi_create(n_gpu, nullptr, &mg) != AIX_OK) {
            fprintf(stderr, "Error: %s\n", aix_multi_last_error(nullptr));
            return 10;
        }
    }
    aix_ctx *ctx = aix_multi_ctx(mg, 0);
    printf("GPUs: %d%s\n", n_gpu, n_gpu > 1 ? (aix_multi_peer_access(mg) ? " (NVLink peer access)" : " (no peer access: ranges move with cudaMemcpyPeer)") : "");
    MappedFile in;
    if (!in.open(argv[1])) {
        fprintf(stderr, "Error: Cannot open input file: %s\n", argv[1]);
        return 1;
    }
    aix_mphf *m = nullptr;
    printf("Loading perfect hash from: %s\n", argv[2]);
    TOOL_CHECK(ctx, aix_mphf_load_pf(ctx, argv[2], &m));
    std::vector<uint64_t> tf(AIX_TOTAL_13MERS);
    aix_count_stats st;
    double t0 = now_s();
    if (aix_count13_multi(mg, m, in.data, in.size, AIX_FMT_DETECT, tf.data(), &st) != AIX_OK) {
        fprintf(stderr, "Error: %s\n", aix_multi_last_error(mg));
        return 10;
    }
    printf("Processing completed in %.0f ms\n", (now_s() - t0) * 1e3);
    uint64_t uniq = 0, total = 0, mx = 0;
    for (uint64_t c : tf)
        if (c) { ++uniq; total += c; if (c > mx) mx = c; }
    printf("\n=== K-mer Counting Statistics ===\n");
    printf("Sequences processed: %llu\n", (unsigned long long)st.sequences);
    printf("Total k-mers processed: %llu\n", (unsigned long long)st.windows);
    printf("Valid k-mers: %llu\n", (unsigned long long)st.valid);
    printf("Invalid k-mers: %llu\n", (unsigned long long)st.invalid);
    printf("Unique k-mers found: %llu / %llu (%g%%)\n", (unsigned long long)uniq, (unsigned long long)AIX_TOTAL_13MERS,
           100.0 * uniq / AIX_TOTAL_13MERS);
    printf("Total k-mer count: %llu\n", (unsigned long long)total);
    printf("Max k-mer frequency: %llu\n", (unsigned long long)mx);
    printf("Average k-mer frequency: %g\n", uniq ? (double)total / uniq : 0.0);
    printf("Saving counts to: %s\n", argv[3]);
    if (!write_file(argv[3], tf.data(), tf.size() * 8)) {
        fprintf(stderr, "Error: Cannot create output file: %s\n", argv[3]);
        return 1;
    }
    printf("Counts saved successfully (%llu MB)\n", (unsigned long long)(tf.size() * 8 / (1024 * 1024)));
    aix_mphf_destroy(ctx, m);
    aix_multi_destroy(mg);
    return 0;
}
