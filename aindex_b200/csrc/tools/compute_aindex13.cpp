// compute_aindex13 <reads_file> <hash_file> <tf_file> <output_prefix> <num_threads>
//                  [_ _ index_bin indices_bin]
// GPU version of the reference tool (src/compute_aindex13.cpp:327-404).  The tf file is the
// 4^13 x uint64 output of count_kmers13 (the reference binary misreads it as uint32,
// SURVEY 2.3#2; a uint32 file of 4^13 entries is accepted too).
#include "tool_common.hpp"

int main(int argc, char **argv) {
    if (argc < 6) {
        fprintf(stderr, "Compute AIndex for 13-mers with perfect hash.\nExpected arguments: %s <reads_file> <hash_file> <tf_file> "
                        "<output_prefix> <num_threads> [pos_bin] [index_bin] [indices_bin]\n", argv[0]);
        return 1;
    }
    const std::string reads_file = argv[1], pf = argv[2], tf_file = argv[3], prefix = argv[4];
    std::string index_bin = prefix + ".index.bin", indices_bin = prefix + ".indices.bin";
    if (argc > 9) {  // compute_aindex13.cpp:344-350
        index_bin = argv[8];
        indices_bin = argv[9];
    }
    aix_ctx *ctx = nullptr;
    if (aix_ctx_create(tool_device(), &ctx) != AIX_OK) {
        fprintf(stderr, "Error: %s\n", aix_last_error(nullptr));
        return 10;
    }
    std::vector<uint8_t> tb;
    if (!read_whole(tf_file, tb)) {
        fprintf(stderr, "Failed to open tf file: %s\n", tf_file.c_str());
        return 1;
    }
    std::vector<uint64_t> tf64(AIX_TOTAL_13MERS);
    if (tb.size() == AIX_TOTAL_13MERS * 8) memcpy(tf64.data(), tb.data(), tb.size());
    else if (tb.size() == AIX_TOTAL_13MERS * 4)
        for (uint64_t i = 0; i < AIX_TOTAL_13MERS; ++i) tf64[i] = ((const uint32_t *)tb.data())[i];
    else {
        fprintf(stderr, "tf file must hold 4^13 uint64 (or uint32) values: %s\n", tf_file.c_str());
        return 1;
    }
    MappedFile reads;
    if (!reads.open(reads_file.c_str())) {
        fprintf(stderr, "Failed to open reads file: %s\n", reads_file.c_str());
        return 1;
    }
    aix_mphf *m = nullptr;
    aix_index13 *ix = nullptr;
    TOOL_CHECK(ctx, aix_mphf_load_pf(ctx, pf.c_str(), &m));
    TOOL_CHECK(ctx, aix_index13_upload(ctx, m, tf64.data(), &ix));
    uint64_t total = 0;
    TOOL_CHECK(ctx, aix_positions_total13(ctx, ix, &total));
    printf("\ttotal_size: %llu\n", (unsigned long long)total);
    std::vector<uint64_t> indices(AIX_TOTAL_13MERS + 1), positions(total);
    TOOL_CHECK(ctx, aix_positions_build13(ctx, ix, reads.data, reads.size, indices.data(), positions.data()));
    if (!write_file(index_bin, positions.data(), positions.size() * 8) || !write_file(indices_bin, indices.data(), indices.size() * 8)) {
        fprintf(stderr, "Cannot write output files\n");
        return 1;
    }
    printf("All files saved successfully.\n");
    aix_index13_destroy(ctx, ix);
    aix_mphf_destroy(ctx, m);
    aix_ctx_destroy(ctx);
    return 0;
}
