// compute_aindex <reads_file> <hash_file> <output_prefix> <num_threads> <k> <tf_file>
//                <kmers_bin_file> <kmers_text_file> [_ index_bin indices_bin]
// GPU version of the reference tool (src/compute_aindex.cpp:28-115): positions index of a
// 23-mer index over a .reads file -> {prefix}.index.bin (positions) + {prefix}.indices.bin.
// num_threads is accepted and ignored; the output equals the reference's 1-thread output.
#include "tool_common.hpp"

int main(int argc, char **argv) {
    if (argc < 9) {
        fprintf(stderr, "Compute AIndex index for genome with pf.\nExpected arguments: %s <reads_file> <hash_file> <output_prefix> "
                        "<num_threads> <k> <tf_file> <kmers_bin_file> <kmers_text_file> [_ index_bin indices_bin]\n", argv[0]);
        return 1;
    }
    const std::string reads_file = argv[1], pf = argv[2], prefix = argv[3], tf_file = argv[6], kmers_bin = argv[7];
    const int k = atoi(argv[5]);
    if (k != 23) {
        fprintf(stderr, "Error: this tool builds the 23-mer positions index (k = %d given); use compute_aindex13 for 13-mers\n", k);
        return 1;
    }
    std::string index_bin = prefix + ".index.bin", indices_bin = prefix + ".indices.bin";
    if (argc > 11) {  // compute_aindex.cpp:59-63: optional outputs are argv[10], argv[11]
        index_bin = argv[10];
        indices_bin = argv[11];
    }
    aix_ctx *ctx = nullptr;
    if (aix_ctx_create(tool_device(), &ctx) != AIX_OK) {
        fprintf(stderr, "Error: %s\n", aix_last_error(nullptr));
        return 10;
    }
    std::vector<uint8_t> kb, tb;
    if (!read_whole(kmers_bin, kb) || !read_whole(tf_file, tb)) {
        fprintf(stderr, "Failed to open kmers/tf file\n");
        return 10;
    }
    const uint64_t n = kb.size() / 8;
    if (tb.size() / 4 < n) {
        fprintf(stderr, "tf file shorter than kmers file\n");
        return 10;
    }
    MappedFile reads;
    if (!reads.open(reads_file.c_str())) {
        fprintf(stderr, "Failed to open reads file: %s\n", reads_file.c_str());
        return 10;
    }
    aix_mphf *m = nullptr;
    aix_index23 *ix = nullptr;
    TOOL_CHECK(ctx, aix_mphf_load_pf(ctx, pf.c_str(), &m));
    TOOL_CHECK(ctx, aix_index23_upload(ctx, m, (const uint64_t *)kb.data(), (const uint32_t *)tb.data(), n, &ix));
    uint64_t total = 0;
    TOOL_CHECK(ctx, aix_positions_total23(ctx, ix, &total));
    printf("\ttotal_size: %llu\n", (unsigned long long)total);
    std::vector<uint64_t> indices(n + 1), positions(total);
    double t0 = now_s();
    TOOL_CHECK(ctx, aix_positions_build23(ctx, ix, reads.data, reads.size, indices.data(), positions.data()));
    printf("Building index... done in %.0f ms\n", (now_s() - t0) * 1e3);
    if (!write_file(index_bin, positions.data(), positions.size() * 8) || !write_file(indices_bin, indices.data(), indices.size() * 8)) {
        fprintf(stderr, "Cannot open file for writting.\n");
        return 12;
    }
    aix_index23_destroy(ctx, ix);
    aix_mphf_destroy(ctx, m);
    aix_ctx_destroy(ctx);
    return 0;
}
