// ref_harness.cpp -- TEST/BENCH INFRASTRUCTURE ONLY (never on the product path).
//
// A thin driver around the UNMODIFIED reference headers (src/hash.hpp, src/kmers.hpp,
// src/emphf/*): it is compiled by oracle/build_ref.sh with -I/root/reference/src and linked
// against the reference's own objects, so every lookup below executes reference code:
//   load_hash            src/hash.cpp:367-450
//   PHASH_MAP::get_freq  src/hash.hpp:123-140, :203-206   (what get_tf_value_23mer does)
//   HASHER::lookup       src/emphf/mphf.hpp:79-89
// It exists because the reference's batch query (AindexWrapper::get_tf_values,
// python_wrapper.cpp:653-664) is single threaded and SURVEY 8(d) asks for "the reference's own
// multithreaded C++ path": the same per-query function called from N std::threads.
//
//   ref_harness tf23  <pf> <tf.bin> <kmers.bin> <queries.bin> <nq> <threads> <out.bin> [reps]
//       queries.bin = nq x 23 ASCII bytes; out.bin = nq x uint32; prints one "seconds=" line per rep
//   ref_harness lookup13 <pf> <threads> <out.bin>
//       mphf lookup of all 4^13 13-mers in numeric order -> uint32 perm
//   ref_harness tf13 <pf> <tf.bin> <threads> <out.bin> <count> [reps]
//       get_tf_value_13mer (python_wrapper.cpp:482-503: length / ACGT check, HASHER::lookup, tf gather,
//       uint32 narrowing) of the 13-mers with 2-bit values 0 .. count-1 -> count x uint32
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <string_view>
#include <thread>
#include <vector>

#include "hash.hpp"
#include "kmers.hpp"
#include "settings.hpp"

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static int cmd_tf23(int argc, char **argv) {
    if (argc < 9) return 2;
    std::string pf = argv[2], tf = argv[3], kb = argv[4], qf = argv[5], of = argv[8];
    uint64_t nq = strtoull(argv[6], nullptr, 10);
    unsigned threads = (unsigned)atoi(argv[7]);
    int reps = argc > 9 ? atoi(argv[9]) : 1;
    if (threads == 0) threads = 1;
    Settings::K = 23;
    PHASH_MAP hm;
    double t0 = now_s();
    std::cout.setstate(std::ios_base::failbit);  // the loader prints progress bars
    load_hash(hm, pf, tf, kb, "");
    std::cout.clear();
    printf("load_seconds=%.3f n=%llu\n", now_s() - t0, (unsigned long long)hm.n);
    std::vector<char> q(nq * 23);
    {
        std::ifstream in(qf, std::ios::binary);
        in.read(q.data(), (std::streamsize)q.size());
        if ((uint64_t)in.gcount() != q.size()) { fprintf(stderr, "short query file\n"); return 3; }
    }
    std::vector<uint32_t> out(nq);
    for (int r = 0; r < reps; ++r) {
        double t1 = now_s();
        std::vector<std::thread> th;
        for (unsigned t = 0; t < threads; ++t) {
            th.emplace_back([&, t]() {
                uint64_t b = nq * t / threads, e = nq * (t + 1) / threads;
                for (uint64_t i = b; i < e; ++i) out[i] = hm.get_freq(std::string_view(q.data() + i * 23, 23));
            });
        }
        for (auto &x : th) x.join();
        printf("seconds=%.6f queries=%llu threads=%u\n", now_s() - t1, (unsigned long long)nq, threads);
        fflush(stdout);
    }
    std::ofstream o(of, std::ios::binary);
    o.write((const char *)out.data(), (std::streamsize)(nq * 4));
    return 0;
}

static int cmd_lookup13(int argc, char **argv) {
    if (argc < 5) return 2;
    HASHER h;
    {
        std::ifstream is(argv[2], std::ios::binary);
        if (!is) return 3;
        h.load(is);
    }
    unsigned threads = (unsigned)atoi(argv[3]);
    if (threads == 0) threads = 1;
    const uint64_t N = 1ull << 26;
    std::vector<uint32_t> out(N);
    double t1 = now_s();
    std::vector<std::thread> th;
    for (unsigned t = 0; t < threads; ++t) {
        th.emplace_back([&, t]() {
            emphf::stl_string_adaptor ad;
            std::string s(13, 'A');
            for (uint64_t v = N * t / threads; v < N * (t + 1) / threads; ++v) {
                get_bitset_dna13((uint32_t)v, s, 13);
                out[v] = (uint32_t)h.lookup(s, ad);
            }
        });
    }
    for (auto &x : th) x.join();
    printf("seconds=%.6f queries=%llu threads=%u\n", now_s() - t1, (unsigned long long)N, threads);
    std::ofstream o(argv[4], std::ios::binary);
    o.write((const char *)out.data(), (std::streamsize)(N * 4));
    return 0;
}

static int cmd_tf13(int argc, char **argv) {
    if (argc < 7) return 2;
    HASHER h;
    {
        std::ifstream is(argv[2], std::ios::binary);
        if (!is) return 3;
        h.load(is);
    }
    const uint64_t N = 1ull << 26;
    std::vector<uint64_t> tf(N);
    {
        std::ifstream in(argv[3], std::ios::binary);
        in.read((char *)tf.data(), (std::streamsize)(N * 8));
        if ((uint64_t)in.gcount() != N * 8) { fprintf(stderr, "short tf file\n"); return 3; }
    }
    unsigned threads = (unsigned)atoi(argv[4]);
    if (threads == 0) threads = 1;
    uint64_t count = strtoull(argv[6], nullptr, 10);
    if (count > N) count = N;
    int reps = argc > 7 ? atoi(argv[7]) : 1;
    std::vector<uint32_t> out(count);
    for (int r = 0; r < reps; ++r) {
        double t1 = now_s();
        std::vector<std::thread> th;
        for (unsigned t = 0; t < threads; ++t) {
            th.emplace_back([&, t]() {
                emphf::stl_string_adaptor ad;
                std::string s(13, 'A');
                for (uint64_t v = count * t / threads; v < count * (t + 1) / threads; ++v) {
                    get_bitset_dna13((uint32_t)v, s, 13);
                    bool ok = s.size() == 13;
                    for (char c : s) ok = ok && (c == 'A' || c == 'C' || c == 'G' || c == 'T');
                    uint64_t id = ok ? h.lookup(s, ad) : N;
                    out[v] = id < N ? (uint32_t)tf[id] : 0u;
                }
            });
        }
        for (auto &x : th) x.join();
        printf("seconds=%.6f queries=%llu threads=%u\n", now_s() - t1, (unsigned long long)count, threads);
        fflush(stdout);
    }
    std::ofstream o(argv[5], std::ios::binary);
    o.write((const char *)out.data(), (std::streamsize)(count * 4));
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: ref_harness tf23|lookup13|tf13 ...\n"); return 2; }
    if (!strcmp(argv[1], "tf13")) return cmd_tf13(argc, argv);
    if (!strcmp(argv[1], "tf23")) return cmd_tf23(argc, argv);
    if (!strcmp(argv[1], "lookup13")) return cmd_lookup13(argc, argv);
    return 2;
}
