// tool_common.hpp -- shared helpers of the GPU command-line tools (same argv contracts as the
// reference binaries; all work goes through the C-ABI of include/aindex_cuda.h).
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <vector>

#include "aindex_cuda.h"

struct MappedFile {
    const uint8_t *data = nullptr;
    size_t size = 0;
    bool open(const char *path) {
        int fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        fstat(fd, &st);
        size = (size_t)st.st_size;
        if (size) {
            void *p = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (p == MAP_FAILED) { ::close(fd); return false; }
            data = (const uint8_t *)p;
        }
        ::close(fd);
        return true;
    }
    ~MappedFile() { if (data) munmap((void *)data, size); }
};

inline bool write_file(const std::string &path, const void *p, size_t bytes) {
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) return false;
    bool ok = bytes == 0 || fwrite(p, 1, bytes, f) == bytes;
    return (fclose(f) == 0) && ok;
}

inline bool read_whole(const std::string &path, std::vector<uint8_t> &out) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    bool ok = out.empty() || fread(out.data(), 1, out.size(), f) == out.size();
    fclose(f);
    return ok;
}

inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

#define TOOL_CHECK(ctx, call)                                                   \
    do {                                                                        \
        int rc__ = (call);                                                      \
        if (rc__ != AIX_OK) {                                                   \
            fprintf(stderr, "Error: %s (code %d)\n", aix_last_error(ctx), rc__); \
            return 10;                                                          \
        }                                                                       \
    } while (0)

inline int tool_device() {
    const char *e = getenv("AINDEX_CUDA_DEVICE");
    return e ? atoi(e) : 0;
}
