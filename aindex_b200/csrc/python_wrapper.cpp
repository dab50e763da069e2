// python_wrapper.cpp -- the pybind11 module `aindex_cpp` (class AindexWrapper) of the
// reference, re-hosted on libaindex_cuda.so.
//
// Same module name, class name, method names, argument order/defaults and return types as
// ad3002/aindex src/python_wrapper.cpp:130-1316 (bindings :1320-2122), so
// aindex/core/aindex.py-style code runs unchanged.  Every query goes through the C-ABI of
// include/aindex_cuda.h -- there is no CPU lookup path in this file; what stays on the host
// is file handling (.reads mmap, .ridx parsing) and list <-> buffer conversion.
//
// Deliberate deviations:
//   * missing files / CUDA failures raise Python exceptions instead of std::terminate()/exit()
//   * get_positions on an absent k-mer returns [] (the reference aborts the process)
//   * load_13mer_aindex also maps the positions file, so get_positions_13mer works
//   * get_reads_se_by_kmer implements the documented meaning (reads containing the k-mer);
//     the reference indexes positions[] with a k-mer id (python_wrapper.cpp:857-870)
// Additions (non-breaking): buffer overloads returning numpy arrays, batched coverage /
// positions, count_kmers13, build_positions, build_index_from_reads.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <fstream>
#include <map>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <tuple>
#include <unistd.h>
#include <unordered_map>
#include <vector>

#include "aindex_cuda.h"

namespace py = pybind11;

namespace {

constexpr uint64_t kTotal13 = AIX_TOTAL_13MERS;

struct FileNotFound : std::runtime_error {
    using std::runtime_error::runtime_error;
};

bool file_exists(const std::string &p) {
    struct stat st;
    return ::stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}

void require_file(const std::string &p, const char *what) {
    if (!file_exists(p)) throw FileNotFound(std::string(what) + " not found: " + p);
}

// read-only mmap of a whole file
struct Mapped {
    void *ptr = nullptr;
    size_t size = 0;
    ~Mapped() { reset(); }
    void reset() {
        if (ptr && size) munmap(ptr, size);
        ptr = nullptr;
        size = 0;
    }
    void open(const std::string &path) {
        reset();
        int fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) throw FileNotFound("cannot open " + path);
        struct stat st;
        fstat(fd, &st);
        size = (size_t)st.st_size;
        if (size) {
            ptr = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (ptr == MAP_FAILED) {
                ptr = nullptr;
                ::close(fd);
                throw std::runtime_error("mmap failed: " + path);
            }
        }
        ::close(fd);
    }
};

std::vector<uint8_t> slurp(const std::string &path) {
    std::ifstream in(path, std::ios::binary | std::ios::ate);
    if (!in) throw FileNotFound("cannot open " + path);
    std::streamsize n = in.tellg();
    in.seekg(0);
    std::vector<uint8_t> buf((size_t)std::max<std::streamsize>(n, 0));
    if (n > 0) in.read((char *)buf.data(), n);
    return buf;
}

// list[str] -> fixed-stride records (+ lengths when not uniform)
struct Records {
    std::vector<uint8_t> bytes;
    std::vector<uint8_t> lens;
    uint32_t stride = 1;
    uint64_t q = 0;
    bool uniform = true;
    const uint8_t *lens_ptr() const { return uniform ? nullptr : lens.data(); }
};

Records pack_records(const std::vector<std::string> &kmers) {
    Records r;
    r.q = kmers.size();
    size_t mx = 1;
    for (auto &s : kmers) mx = std::max(mx, s.size());
    if (mx > 255) throw std::invalid_argument("query strings longer than 255 characters are not supported");
    r.stride = (uint32_t)mx;
    for (auto &s : kmers)
        if (s.size() != mx) r.uniform = false;
    r.bytes.assign((size_t)r.q * r.stride, 0);
    if (!r.uniform) r.lens.resize(r.q);
    for (uint64_t i = 0; i < r.q; ++i) {
        memcpy(r.bytes.data() + i * r.stride, kmers[i].data(), kmers[i].size());
        if (!r.uniform) r.lens[i] = (uint8_t)kmers[i].size();
    }
    return r;
}

// list of str (or bytes) -> records without building a std::vector<std::string> first: the UTF-8 view of an
// ASCII str is the object's own buffer, so a 23-mer costs one memcpy instead of a heap allocation and two copies.
// Must be called with the GIL held.
Records pack_records_py(const py::list &kmers) {
    Records r;
    r.q = (uint64_t)kmers.size();
    std::vector<const char *> ptr(r.q);
    std::vector<Py_ssize_t> len(r.q);
    size_t mx = 1;
    for (uint64_t i = 0; i < r.q; ++i) {
        PyObject *o = PyList_GET_ITEM(kmers.ptr(), (Py_ssize_t)i);
        if (PyUnicode_Check(o)) {
            ptr[i] = PyUnicode_AsUTF8AndSize(o, &len[i]);
            if (!ptr[i]) throw py::error_already_set();
        } else if (PyBytes_Check(o)) {
            char *b = nullptr;
            if (PyBytes_AsStringAndSize(o, &b, &len[i]) != 0) throw py::error_already_set();
            ptr[i] = b;
        } else {
            throw py::type_error("k-mers must be str or bytes");
        }
        mx = std::max(mx, (size_t)len[i]);
    }
    if (mx > 255) throw std::invalid_argument("query strings longer than 255 characters are not supported");
    r.stride = (uint32_t)mx;
    for (uint64_t i = 0; i < r.q; ++i)
        if ((size_t)len[i] != mx) { r.uniform = false; break; }
    r.bytes.assign((size_t)r.q * r.stride, 0);
    if (!r.uniform) r.lens.resize(r.q);
    for (uint64_t i = 0; i < r.q; ++i) {
        memcpy(r.bytes.data() + i * r.stride, ptr[i], (size_t)len[i]);
        if (!r.uniform) r.lens[i] = (uint8_t)len[i];
    }
    return r;
}

struct Interval {  // python_wrapper.cpp:44-53 (end is stored as end+1, :271)
    uint64_t rid, start, end;
};

}  // namespace

class AindexWrapper {
    aix_ctx *ctx = nullptr;
    mutable std::recursive_mutex mu;  // taken while the GIL is released around the batch entry points
    aix_mphf *mphf23 = nullptr;
    aix_index23 *ix23 = nullptr;
    aix_positions *pos23 = nullptr;
    std::vector<uint64_t> checker;  // host copy for kid -> k-mer (get_kmer_by_kid, get_kmer_info)
    std::vector<uint32_t> tf23;
    uint32_t max_tf = 0;

    bool is_13mer_mode = false;
    aix_mphf *mphf13 = nullptr;
    aix_index13 *ix13 = nullptr;
    aix_positions *pos13 = nullptr;
    Mapped tf13_map;  // the 4^13 x u64 tf file (get_13mer_tf_array & co)

    Mapped reads_map;
    std::vector<char> reads_mem;
    const char *reads = nullptr;
    std::vector<uint64_t> start_positions;
    std::unordered_map<uint64_t, uint64_t> start2end;
    std::vector<Interval> intervals;
    bool intervals_sorted = true;

public:
    bool aindex_loaded = false;
    uint64_t n_reads = 0;
    uint64_t n_kmers = 0;
    uint64_t reads_size = 0;

    AindexWrapper() = default;
    AindexWrapper(const AindexWrapper &) = delete;
    AindexWrapper &operator=(const AindexWrapper &) = delete;

    ~AindexWrapper() {
        if (ctx) {
            aix_positions_destroy(ctx, pos23);
            aix_positions_destroy(ctx, pos13);
            aix_index23_destroy(ctx, ix23);
            aix_index13_destroy(ctx, ix13);
            aix_mphf_destroy(ctx, mphf23);
            aix_mphf_destroy(ctx, mphf13);
            aix_ctx_destroy(ctx);
        }
    }

private:
    void ensure_ctx() {
        if (ctx) return;
        int dev = 0;
        if (const char *e = getenv("AINDEX_CUDA_DEVICE")) dev = atoi(e);
        else if (const char *l = getenv("LOCAL_RANK")) dev = atoi(l);
        int rc = aix_ctx_create(dev, &ctx);
        if (rc != AIX_OK) {
            ctx = nullptr;
            throw std::runtime_error(std::string("aindex_cpp needs a CUDA device: ") + aix_last_error(nullptr));
        }
    }
    void check(int rc) const {
        if (rc == AIX_OK) return;
        std::string msg = aix_last_error(ctx);
        if (rc == AIX_ERR_IO) throw FileNotFound(msg);
        if (rc == AIX_ERR_ARG) throw std::invalid_argument(msg);
        throw std::runtime_error(msg);
    }
    // a C-ABI call made with the GIL held: still exclusive with a batch call another thread runs with the GIL released
    template <typename F>
    void locked(F &&f) const {
        int rc;
        {
            std::lock_guard<std::recursive_mutex> device_lock(mu);
            rc = f();
        }
        check(rc);
    }
    void require23() const {
        if (!ix23) throw std::runtime_error("23-mer index not loaded");
    }

    template <typename T>
    std::vector<T> run23(const Records &r, int mode, size_t per = 1) const {
        require23();
        std::vector<T> out(r.q * per);
        if (r.q) {
            py::gil_scoped_release nogil;
            std::lock_guard<std::recursive_mutex> device_lock(mu);  // an aix_ctx serves one caller at a time (the reference relied on the GIL)
            check(aix_tf23_batch(ctx, ix23, r.bytes.data(), r.stride, r.lens_ptr(), r.q, mode, out.data()));
        }
        return out;
    }
    template <typename T>
    std::vector<T> run13(const Records &r, int mode, size_t per = 1) const {
        std::vector<T> out(r.q * per);
        if (r.q) {
            py::gil_scoped_release nogil;
            std::lock_guard<std::recursive_mutex> device_lock(mu);  // an aix_ctx serves one caller at a time (the reference relied on the GIL)
            check(aix_tf13_batch(ctx, ix13, r.bytes.data(), r.stride, r.lens_ptr(), r.q, mode, out.data()));
        }
        return out;
    }
    template <typename T>
    std::vector<T> query23(const std::vector<std::string> &kmers, int mode, size_t per = 1) const {
        return run23<T>(pack_records(kmers), mode, per);
    }
    template <typename T>
    std::vector<T> query13(const std::vector<std::string> &kmers, int mode, size_t per = 1) const {
        return run13<T>(pack_records(kmers), mode, per);
    }

public:
    // the batch form of get_tf_values: list in, list out, no per-string C++ objects in between
    py::list tf_values_list(const py::list &kmers) const {
        Records r = pack_records_py(kmers);
        std::vector<uint32_t> out(r.q);
        if (r.q) {
            py::gil_scoped_release nogil;
            std::lock_guard<std::recursive_mutex> device_lock(mu);  // an aix_ctx serves one caller at a time (the reference relied on the GIL)
            if (is_13mer_mode) check(aix_tf13_batch(ctx, ix13, r.bytes.data(), r.stride, r.lens_ptr(), r.q, AIX_Q_TF, out.data()));
            else {
                if (!ix23) throw std::runtime_error("23-mer index not loaded");
                locked([&] { return aix_tf23_batch(ctx, ix23, r.bytes.data(), r.stride, r.lens_ptr(), r.q, AIX_Q_TF, out.data()); });
            }
        }
        py::list res(r.q);
        for (uint64_t i = 0; i < r.q; ++i) PyList_SET_ITEM(res.ptr(), (Py_ssize_t)i, PyLong_FromUnsignedLong(out[i]));
        return res;
    }

    // ------------------------------------------------------------------ loaders
    // load / load_hash_file (python_wrapper.cpp:228-259) -> load_hash (hash.cpp:367-450)
    void load(std::string hash_filename, std::string tf_file, std::string kmers_bin_filename,
              std::string /*kmers_text_filename*/) {
        require_file(hash_filename, "hash file");
        require_file(tf_file, "tf file");
        require_file(kmers_bin_filename, "kmers_bin file");
        ensure_ctx();
        std::vector<uint8_t> kb = slurp(kmers_bin_filename), tb = slurp(tf_file);
        uint64_t n = kb.size() / 8;  // hash.cpp:388-392
        if (tb.size() / 4 < n) throw std::runtime_error("tf file shorter than kmers file");
        std::lock_guard<std::recursive_mutex> device_lock(mu);  // teardown + upload are one critical section: a batch call on another thread must not see freed objects
        is_13mer_mode = false;
        n_kmers = 0;
        aix_positions_destroy(ctx, pos23); pos23 = nullptr;
        aix_index23_destroy(ctx, ix23); ix23 = nullptr;
        aix_mphf_destroy(ctx, mphf23); mphf23 = nullptr;
        locked([&] { return aix_mphf_load_pf(ctx, hash_filename.c_str(), &mphf23); });
        checker.assign((const uint64_t *)kb.data(), (const uint64_t *)kb.data() + n);
        tf23.assign((const uint32_t *)tb.data(), (const uint32_t *)tb.data() + n);
        locked([&] { return aix_index23_upload(ctx, mphf23, checker.data(), tf23.data(), n, &ix23); });
        n_kmers = n;
        is_13mer_mode = false;
    }
    void load_hash_file(std::string a, std::string b, std::string c, std::string d) { load(a, b, c, d); }

    // load_reads_index (python_wrapper.cpp:261-279)
    void load_reads_index(const std::string &index_file) {
        std::ifstream fin(index_file);
        if (!fin.is_open()) throw FileNotFound("Error opening index file: " + index_file);
        n_reads = 0;
        uint64_t rid, s, e, prev = 0;
        while (fin >> rid >> s >> e) {
            intervals.push_back({rid, s, e + 1});
            if (s < prev) intervals_sorted = false;
            prev = s;
            start_positions.push_back(s);
            start2end[s] = e;
            n_reads++;
        }
    }

    void load_reads(std::string reads_file) {  // :281-322 (mmap)
        require_file(reads_file, "reads file");
        reads_mem.clear();
        reads_map.open(reads_file);
        reads = (const char *)reads_map.ptr;
        reads_size = reads_map.size;
        std::string ridx = reads_file.substr(0, reads_file.find_last_of(".")) + ".ridx";
        load_reads_index(ridx);
    }

    void load_reads_in_memory(std::string reads_file) {  // :324-359
        require_file(reads_file, "reads file");
        reads_map.reset();
        std::vector<uint8_t> b = slurp(reads_file);
        reads_mem.assign(b.begin(), b.end());
        reads = reads_mem.data();
        reads_size = reads_mem.size();
        std::string ridx = reads_file.substr(0, reads_file.find_last_of(".")) + ".ridx";
        load_reads_index(ridx);
    }

    // load_aindex (:361-402): .index.bin = positions, .indices.bin = offsets; both go to HBM
    void load_aindex(std::string index_file, std::string indices_file, uint32_t _max_tf) {
        require23();
        require_file(index_file, "index file");
        require_file(indices_file, "indices file");
        max_tf = _max_tf;
        Mapped pos, ind;
        pos.open(index_file);
        ind.open(indices_file);
        std::lock_guard<std::recursive_mutex> device_lock(mu);  // teardown + upload are one critical section: a batch call on another thread must not see freed objects
        aindex_loaded = false;
        aix_positions_destroy(ctx, pos23); pos23 = nullptr;
        locked([&] { return aix_positions_upload(ctx, (const uint64_t *)ind.ptr, ind.size / 8, (const uint64_t *)pos.ptr, pos.size / 8, &pos23); });
        aindex_loaded = true;
    }

    void load_13mer_index(const std::string &hash_file, const std::string &tf_file) {  // :404-437
        require_file(hash_file, "13-mer hash file");
        require_file(tf_file, "13-mer tf file");
        ensure_ctx();
        tf13_map.open(tf_file);
        if (tf13_map.size < kTotal13 * 8) throw std::runtime_error("13-mer tf file must hold 4^13 uint64 values: " + tf_file);
        std::lock_guard<std::recursive_mutex> device_lock(mu);  // teardown + upload are one critical section: a batch call on another thread must not see freed objects
        is_13mer_mode = false;  // stays false if the reload fails: run13 must never see a null ix13 in 13-mer mode
        n_kmers = 0;
        aix_positions_destroy(ctx, pos13); pos13 = nullptr;
        aix_index13_destroy(ctx, ix13); ix13 = nullptr;
        aix_mphf_destroy(ctx, mphf13); mphf13 = nullptr;
        locked([&] { return aix_mphf_load_pf(ctx, hash_file.c_str(), &mphf13); });
        locked([&] { return aix_index13_upload(ctx, mphf13, (const uint64_t *)tf13_map.ptr, &ix13); });
        is_13mer_mode = true;
        n_kmers = kTotal13;
    }

    void load_13mer_aindex(const std::string &index_file, const std::string &indices_file) {  // :439-471
        if (!ix13) throw std::runtime_error("13-mer index not loaded");
        require_file(index_file, "13-mer index file");
        require_file(indices_file, "13-mer indices file");
        Mapped pos, ind;
        pos.open(index_file);
        ind.open(indices_file);
        std::lock_guard<std::recursive_mutex> device_lock(mu);  // teardown + upload are one critical section: a batch call on another thread must not see freed objects
        aix_positions_destroy(ctx, pos13); pos13 = nullptr;
        locked([&] { return aix_positions_upload(ctx, (const uint64_t *)ind.ptr, ind.size / 8, (const uint64_t *)pos.ptr, pos.size / 8, &pos13); });
        aindex_loaded = true;
    }

    void load_from_prefix_23mer(const std::string &prefix, const std::string &reads_file = "") {  // :1103-1132
        load(prefix + ".pf", prefix + ".tf.bin", prefix + ".kmers.bin", prefix + ".txt");
        if (!reads_file.empty()) load_reads(reads_file);
    }
    void load_aindex_from_prefix_23mer(const std::string &prefix, uint32_t mtf, const std::string &reads_file = "") {  // :1134-1160
        load_aindex(prefix + ".index.bin", prefix + ".indices.bin", mtf);
        if (!reads_file.empty() && reads == nullptr) load_reads(reads_file);
    }
    void load_from_prefix_13mer(const std::string &prefix, const std::string &reads_file = "") {  // :1162-1188
        load_13mer_index(prefix + ".pf", prefix + ".tf.bin");
        if (!reads_file.empty()) load_reads(reads_file);
    }
    void load_aindex_from_prefix_13mer(const std::string &prefix, const std::string &reads_file = "") {  // :1190-1216
        load_13mer_aindex(prefix + ".index.bin", prefix + ".indices.bin");
        if (!reads_file.empty() && reads == nullptr) load_reads(reads_file);
    }

    // ------------------------------------------------------------------ tf queries
    // list overloads of the batch calls: the str buffers are read in place (pack_records_py), registered in front of
    // the std::vector<std::string> forms, which stay for other sequences
    std::vector<uint32_t> tf_values_23mer_list(const py::list &kmers) { return run23<uint32_t>(pack_records_py(kmers), AIX_Q_TF); }
    std::vector<uint64_t> total_tf_values_23mer_list(const py::list &kmers) { return run23<uint64_t>(pack_records_py(kmers), AIX_Q_TOTAL); }
    std::vector<uint32_t> tf_values_13mer_list(const py::list &kmers) {
        if (!is_13mer_mode) return std::vector<uint32_t>(kmers.size(), 0);
        return run13<uint32_t>(pack_records_py(kmers), AIX_Q_TF);
    }
    std::vector<uint64_t> total_tf_values_13mer_list(const py::list &kmers) {
        if (!is_13mer_mode) return std::vector<uint64_t>(kmers.size(), 0);
        return run13<uint64_t>(pack_records_py(kmers), AIX_Q_TOTAL);
    }
    std::vector<uint32_t> get_tf_values_23mer(const std::vector<std::string> &kmers) { return query23<uint32_t>(kmers, AIX_Q_TF); }
    uint32_t get_tf_value_23mer(const std::string &kmer) { return query23<uint32_t>({kmer}, AIX_Q_TF)[0]; }

    std::vector<uint32_t> get_tf_values_13mer(const std::vector<std::string> &kmers) {  // :938-980
        if (!is_13mer_mode) return std::vector<uint32_t>(kmers.size(), 0);
        return query13<uint32_t>(kmers, AIX_Q_TF);
    }
    uint32_t get_tf_value_13mer(const std::string &kmer) { return get_tf_values_13mer({kmer})[0]; }

    uint32_t get_tf_value(const std::string &kmer) {  // :644-650
        return is_13mer_mode ? get_tf_value_13mer(kmer) : get_tf_value_23mer(kmer);
    }
    std::vector<uint32_t> get_tf_values(const std::vector<std::string> &kmers) {  // :653-664
        return is_13mer_mode ? get_tf_values_13mer(kmers) : get_tf_values_23mer(kmers);
    }

    // buffer overload: uint8[q, k] (or bytes of q*k characters) -> uint32[q], no per-string objects
    py::array_t<uint32_t> get_tf_values_array(py::array_t<uint8_t, py::array::c_style | py::array::forcecast> recs) {
        if (recs.ndim() != 2) throw std::invalid_argument("expected a uint8 array of shape (q, k)");
        uint64_t q = (uint64_t)recs.shape(0);
        uint32_t stride = (uint32_t)recs.shape(1);
        py::array_t<uint32_t> out((py::ssize_t)q);
        if (q) {
            const uint8_t *in = recs.data();
            uint32_t *o = out.mutable_data();
            if (is_13mer_mode) {
                py::gil_scoped_release nogil;
            std::lock_guard<std::recursive_mutex> device_lock(mu);  // an aix_ctx serves one caller at a time (the reference relied on the GIL)
                check(aix_tf13_batch(ctx, ix13, in, stride, nullptr, q, AIX_Q_TF, o));
            } else {
                require23();
                py::gil_scoped_release nogil;
            std::lock_guard<std::recursive_mutex> device_lock(mu);  // an aix_ctx serves one caller at a time (the reference relied on the GIL)
                check(aix_tf23_batch(ctx, ix23, in, stride, nullptr, q, AIX_Q_TF, o));
            }
        }
        return out;
    }

    uint64_t get_total_tf_value_13mer(const std::string &kmer) {  // :522-545
        if (!is_13mer_mode) return 0;
        return query13<uint64_t>({kmer}, AIX_Q_TOTAL)[0];
    }
    std::vector<uint64_t> get_total_tf_values_13mer(const std::vector<std::string> &kmers) {  // :550-565
        if (!is_13mer_mode) return std::vector<uint64_t>(kmers.size(), 0);
        return query13<uint64_t>(kmers, AIX_Q_TOTAL);
    }
    std::pair<uint64_t, uint64_t> get_tf_both_directions_13mer(const std::string &kmer) {  // :570-591
        if (!is_13mer_mode) return {0, 0};
        auto v = query13<uint64_t>({kmer}, AIX_Q_BOTH, 2);
        return {v[0], v[1]};
    }
    std::vector<std::pair<uint64_t, uint64_t>> get_tf_both_directions_13mer_batch(const std::vector<std::string> &kmers) {  // :597-608
        std::vector<std::pair<uint64_t, uint64_t>> out(kmers.size(), {0, 0});
        if (!is_13mer_mode) return out;
        auto v = query13<uint64_t>(kmers, AIX_Q_BOTH, 2);
        for (size_t i = 0; i < kmers.size(); ++i) out[i] = {v[2 * i], v[2 * i + 1]};
        return out;
    }
    std::string get_reverse_complement_13mer(const std::string &kmer) {  // :505-517 (string op, any length)
        std::string rc(kmer.rbegin(), kmer.rend());
        for (char &c : rc) {
            switch (c) {
                case 'A': c = 'T'; break;
                case 'T': c = 'A'; break;
                case 'G': c = 'C'; break;
                case 'C': c = 'G'; break;
            }
        }
        return rc;
    }

    uint64_t get_total_tf_value_23mer(const std::string &kmer) { return query23<uint64_t>({kmer}, AIX_Q_TOTAL)[0]; }  // :1230-1246
    std::vector<uint64_t> get_total_tf_values_23mer(const std::vector<std::string> &kmers) { return query23<uint64_t>(kmers, AIX_Q_TOTAL); }
    std::pair<uint32_t, uint32_t> get_tf_both_directions_23mer(const std::string &kmer) {  // :1260-1275
        auto v = query23<uint32_t>({kmer}, AIX_Q_BOTH, 2);
        return {v[0], v[1]};
    }
    std::vector<std::pair<uint32_t, uint32_t>> get_tf_both_directions_23mer_batch(const std::vector<std::string> &kmers) {
        auto v = query23<uint32_t>(kmers, AIX_Q_BOTH, 2);
        std::vector<std::pair<uint32_t, uint32_t>> out(kmers.size());
        for (size_t i = 0; i < kmers.size(); ++i) out[i] = {v[2 * i], v[2 * i + 1]};
        return out;
    }
    std::string get_reverse_complement_23mer(const std::string &kmer) {  // :1288-1299 (encode -> reverseDNA -> decode)
        if (kmer.length() != 23) return "";
        ensure_ctx();
        uint64_t u = 0, r = 0;
        uint8_t out[23];
        locked([&] { return aix_encode_kmers(ctx, (const uint8_t *)kmer.data(), 23, nullptr, 1, 23, &u); });
        locked([&] { return aix_revcomp_kmers(ctx, &u, 1, 23, &r); });
        locked([&] { return aix_decode_kmers(ctx, &r, 1, 23, out); });
        return std::string((const char *)out, 23);
    }

    // ------------------------------------------------------------------ ids
    std::vector<uint64_t> get_hash_values(std::vector<std::string> kmers) {  // :629-636
        require23();
        Records r = pack_records(kmers);
        std::vector<uint64_t> out(r.q);
        if (r.q) locked([&] { return aix_mphf_lookup(ctx, mphf23, r.bytes.data(), r.stride, r.lens_ptr(), r.q, out.data()); });
        return out;
    }
    uint64_t get_hash_value(std::string kmer) { return get_hash_values({kmer})[0]; }
    uint64_t get_kid_by_kmer(std::string kmer) { return query23<uint64_t>({kmer}, AIX_Q_KID)[0]; }   // :700-716
    uint64_t get_strand(std::string kmer) { return query23<uint64_t>({kmer}, AIX_Q_STRAND)[0]; }    // :726-742

    std::string decode23(uint64_t u) const {
        uint8_t out[23];
        locked([&] { return aix_decode_kmers(ctx, &u, 1, 23, out); });
        return std::string((const char *)out, 23);
    }
    std::string get_kmer_by_kid(uint64_t kid) {  // :718-724
        require23();
        if (kid >= checker.size()) return "";
        return decode23(checker[kid]);
    }
    std::tuple<uint64_t, std::string, std::string> get_kmer_info(uint64_t kid) {  // :744-755
        require23();
        if (kid >= checker.size()) return std::make_tuple((uint64_t)0, std::string(""), std::string(""));
        uint64_t u = checker[kid], r = 0;
        locked([&] { return aix_revcomp_kmers(ctx, &u, 1, 23, &r); });
        return std::make_tuple((uint64_t)tf23[kid], decode23(u), decode23(r));
    }

    // ------------------------------------------------------------------ reads
    // IntervalTree::query (:64-74) returns the first stored interval with
    // start <= pos+1 && end >= pos (end is stored +1): off-by-one quirk kept
    const Interval *find_interval(uint64_t lo, uint64_t hi) const {
        if (intervals_sorted) {
            // ends are increasing too for a .ridx file: first interval with end >= lo
            size_t a = 0, b = intervals.size();
            while (a < b) {
                size_t m = (a + b) / 2;
                if (intervals[m].end >= lo) b = m;
                else a = m + 1;
            }
            if (a < intervals.size() && intervals[a].start <= hi) return &intervals[a];
            return nullptr;
        }
        for (auto &iv : intervals)
            if (iv.start <= hi && iv.end >= lo) return &iv;
        return nullptr;
    }
    uint64_t get_rid(uint64_t pos) {  // :757-772
        if (!aindex_loaded || intervals.empty()) return 0;
        const Interval *iv = find_interval(pos, pos + 1);
        return iv ? iv->rid : 0;
    }
    uint64_t get_start(uint64_t pos) {  // :774-789
        if (!aindex_loaded || intervals.empty()) return 0;
        const Interval *iv = find_interval(pos, pos + 1);
        return iv ? iv->start : 0;
    }
    std::string get_read_by_rid(uint64_t rid) {  // :666-675
        if (start_positions.size() <= rid || !reads) return "";
        uint64_t s = start_positions[rid], e = start2end[s];
        if (e > reads_size || s > e) return "";
        return std::string(reads + s, e - s);
    }
    // batched forms (additions): the same lookups for whole arrays, GIL released, no per-read Python objects
    py::tuple get_reads_by_rids(py::array_t<uint64_t, py::array::c_style | py::array::forcecast> rids) {
        const uint64_t q = (uint64_t)rids.size();
        const uint64_t *r = rids.data();
        py::array_t<uint64_t> offs((py::ssize_t)q + 1);
        uint64_t *o = offs.mutable_data();
        o[0] = 0;
        std::vector<std::pair<uint64_t, uint64_t>> span(q);
        for (uint64_t i = 0; i < q; ++i) {  // get_read_by_rid (:666-675): "" for an unknown rid
            uint64_t s = 0, e = 0;
            if (r[i] < start_positions.size() && reads) {
                s = start_positions[r[i]];
                auto it = start2end.find(s);
                e = it == start2end.end() ? s : it->second;
                if (e > reads_size || s > e) s = e = 0;
            }
            span[i] = {s, e};
            o[i + 1] = o[i] + (e - s);
        }
        std::string out((size_t)o[q], '\0');
        {
            py::gil_scoped_release nogil;
            for (uint64_t i = 0; i < q; ++i)
                if (span[i].second > span[i].first) memcpy(&out[(size_t)o[i]], reads + span[i].first, (size_t)(span[i].second - span[i].first));
        }
        return py::make_tuple(py::bytes(out), offs);
    }
    py::tuple get_rids_and_starts(py::array_t<uint64_t, py::array::c_style | py::array::forcecast> positions) {
        const uint64_t q = (uint64_t)positions.size();
        const uint64_t *p = positions.data();
        py::array_t<uint64_t> rid((py::ssize_t)q), st((py::ssize_t)q);
        uint64_t *pr = rid.mutable_data(), *ps = st.mutable_data();
        const bool ok = aindex_loaded && !intervals.empty();
        {
            py::gil_scoped_release nogil;
            for (uint64_t i = 0; i < q; ++i) {  // get_rid / get_start (:757-789), quirk 2.3#10 included
                const Interval *iv = ok ? find_interval(p[i], p[i] + 1) : nullptr;
                pr[i] = iv ? iv->rid : 0;
                ps[i] = iv ? iv->start : 0;
            }
        }
        return py::make_tuple(rid, st);
    }
    std::string get_read(uint64_t start, uint64_t end, bool revcomp = false) {  // :677-698
        if (!reads || start >= reads_size || end >= reads_size || start > end) return "";
        std::string read(reads + start, end - start);
        if (!revcomp) return read;
        std::string rev;
        rev.reserve(read.size());
        for (size_t i = read.size(); i-- > 0;) {
            switch (read[i]) {
                case 'A': rev += 'T'; break;
                case 'T': rev += 'A'; break;
                case 'C': rev += 'G'; break;
                case 'G': rev += 'C'; break;
                default: rev += read[i]; break;
            }
        }
        return rev;
    }

    // ------------------------------------------------------------------ positions
    std::vector<uint64_t> positions_of(const std::string &kmer, int k) const {
        std::vector<uint64_t> out;
        const aix_positions *p = k == 23 ? pos23 : pos13;
        if (!p || (k == 23 && !ix23) || (k == 13 && !ix13)) return out;
        uint64_t cnt = 0, offs[2] = {0, 0};
        locked([&] { return aix_positions_query(ctx, ix23, ix13, p, (const uint8_t *)kmer.data(), (uint32_t)kmer.size(), nullptr, 1, k, &cnt, nullptr, nullptr); });
        if (!cnt) return out;
        out.resize(cnt);
        offs[1] = cnt;
        locked([&] { return aix_positions_query(ctx, ix23, ix13, p, (const uint8_t *)kmer.data(), (uint32_t)kmer.size(), nullptr, 1, k, nullptr, offs, out.data()); });
        return out;
    }
    std::vector<uint64_t> get_positions_23mer(const std::string &kmer) { return kmer.size() == 23 ? positions_of(kmer, 23) : std::vector<uint64_t>{}; }  // :800-822
    std::vector<uint64_t> get_positions_13mer(const std::string &kmer) {  // :1070-1101
        if (!is_13mer_mode || kmer.size() != 13) return {};
        return positions_of(kmer, 13);
    }
    std::vector<uint64_t> get_positions(const std::string &kmer) {  // :826-831
        if (kmer.size() == 13) return get_positions_13mer(kmer);
        if (kmer.size() == 23) return get_positions_23mer(kmer);
        return {};
    }
    // batched: -> (offsets uint64[q+1], positions uint64[total])
    py::tuple get_positions_batch(const std::vector<std::string> &kmers, int k) {
        const aix_positions *p = k == 23 ? pos23 : pos13;
        if (!p) throw std::runtime_error("positions index not loaded");
        Records r = pack_records(kmers);
        py::array_t<uint64_t> offs((py::ssize_t)r.q + 1);
        uint64_t *o = offs.mutable_data();
        std::vector<uint64_t> counts(r.q);
        o[0] = 0;
        if (r.q) locked([&] { return aix_positions_query(ctx, ix23, ix13, p, r.bytes.data(), r.stride, r.lens_ptr(), r.q, k, counts.data(), nullptr, nullptr); });
        for (uint64_t i = 0; i < r.q; ++i) o[i + 1] = o[i] + counts[i];
        py::array_t<uint64_t> vals((py::ssize_t)o[r.q]);
        if (o[r.q]) locked([&] { return aix_positions_query(ctx, ix23, ix13, p, r.bytes.data(), r.stride, r.lens_ptr(), r.q, k, nullptr, o, vals.mutable_data()); });
        return py::make_tuple(offs, vals);
    }

    std::vector<std::string> get_reads_se_by_kmer(std::string kmer, uint64_t max_reads) {
        std::vector<std::string> result;
        if (!aindex_loaded || !reads) return result;
        std::vector<uint64_t> last_rid;
        for (uint64_t pos : get_positions(kmer)) {
            if (result.size() >= max_reads) break;
            const Interval *iv = find_interval(pos, pos + kmer.size() - 1);
            if (!iv) continue;
            if (std::find(last_rid.begin(), last_rid.end(), iv->rid) != last_rid.end()) continue;
            last_rid.push_back(iv->rid);
            std::string read = get_read_by_rid(iv->rid);
            if (!read.empty()) result.push_back(read);
        }
        return result;
    }

    // ------------------------------------------------------------------ coverage (aindex.py:314-322)
    py::array_t<uint32_t> get_sequence_coverage(const std::string &seq, uint32_t cutoff = 0, int k = 23) {
        int64_t offs[2] = {0, (int64_t)seq.size()};
        size_t n = seq.size() >= (size_t)k ? seq.size() - k + 1 : 0;
        py::array_t<uint32_t> out((py::ssize_t)n);
        if (n) {
            uint32_t *o = out.mutable_data();
            py::gil_scoped_release nogil;
            std::lock_guard<std::recursive_mutex> device_lock(mu);  // an aix_ctx serves one caller at a time (the reference relied on the GIL)
            check(aix_coverage(ctx, ix23, ix13, (const uint8_t *)seq.data(), offs, 1, k, cutoff, o));
        }
        return out;
    }
    py::array_t<uint32_t> get_sequence_coverage_batch(py::bytes seqs, py::array_t<int64_t, py::array::c_style | py::array::forcecast> offs,
                                                      uint32_t cutoff = 0, int k = 23) {
        std::string_view sv = seqs;
        if (offs.ndim() != 1 || offs.shape(0) < 1) throw std::invalid_argument("offsets must be int64[n_seq + 1]");
        uint64_t n_seq = (uint64_t)offs.shape(0) - 1;
        const int64_t *op = offs.data();
        if (n_seq && (op[0] < 0 || (uint64_t)op[n_seq] > sv.size())) throw std::invalid_argument("offsets out of range");
        uint64_t total = 0;
        for (uint64_t s = 0; s < n_seq; ++s) {
            int64_t len = op[s + 1] - op[s];
            if (len >= k) total += (uint64_t)(len - k + 1);
        }
        py::array_t<uint32_t> out((py::ssize_t)total);
        if (total) {
            uint32_t *o = out.mutable_data();
            py::gil_scoped_release nogil;
            std::lock_guard<std::recursive_mutex> device_lock(mu);  // an aix_ctx serves one caller at a time (the reference relied on the GIL)
            check(aix_coverage(ctx, ix23, ix13, (const uint8_t *)sv.data(), op, n_seq, k, cutoff, o));
        }
        return out;
    }

    // ------------------------------------------------------------------ metadata
    uint64_t get_hash_size() { return is_13mer_mode ? kTotal13 : (uint64_t)checker.size(); }  // :846-851
    uint64_t get_reads_size() { return reads_size; }

    // tf of every 13-mer indexed by its 2-bit value (addition): what get_13mer_tf_array would be if the MPHF were the identity
    py::array_t<uint64_t> get_13mer_tf_array_direct() {
        if (!is_13mer_mode || !ix13) throw std::runtime_error("13-mer index not loaded");
        py::array_t<uint64_t> out((py::ssize_t)kTotal13);
        uint64_t *o = out.mutable_data();
        {
            py::gil_scoped_release nogil;
            std::lock_guard<std::recursive_mutex> device_lock(mu);  // an aix_ctx serves one caller at a time (the reference relied on the GIL)
            check(aix_index13_tf_direct(ctx, ix13, o));
        }
        return out;
    }
    // tf of every stored 23-mer by kid (addition; the host copy the loader keeps for get_kmer_info)
    py::array_t<uint32_t> get_tf_array_23mer() {
        require23();
        py::array_t<uint32_t> out((py::ssize_t)tf23.size());
        if (!tf23.empty()) memcpy(out.mutable_data(), tf23.data(), tf23.size() * 4);
        return out;
    }
    std::vector<uint32_t> get_13mer_tf_array() {  // :983-990 (u64 -> u32 narrowing as in the reference)
        if (!is_13mer_mode) return {};
        const uint64_t *t = (const uint64_t *)tf13_map.ptr;
        return std::vector<uint32_t>(t, t + kTotal13);
    }
    uint32_t get_tf_by_index_13mer(uint64_t index) {  // :993-998
        if (!is_13mer_mode || index >= kTotal13) return 0;
        return (uint32_t)((const uint64_t *)tf13_map.ptr)[index];
    }
    std::map<std::string, uint64_t> get_13mer_statistics() {  // :1037-1068
        std::map<std::string, uint64_t> stats;
        if (!is_13mer_mode) return stats;
        const uint64_t *t = (const uint64_t *)tf13_map.ptr;
        uint64_t nz = 0, mx = 0, tot = 0;
        for (uint64_t i = 0; i < kTotal13; ++i) {
            if (t[i]) { ++nz; tot += t[i]; mx = std::max(mx, t[i]); }
        }
        stats["total_kmers"] = kTotal13;
        stats["non_zero_kmers"] = nz;
        stats["max_frequency"] = mx;
        stats["total_count"] = tot;
        return stats;
    }
    std::string get_index_info() {  // :1001-1034
        std::string info = "Index Info:\n";
        if (is_13mer_mode && tf13_map.ptr) {
            auto st = get_13mer_statistics();
            info += "Mode: 13-mer\n";
            info += "Total k-mers: " + std::to_string(kTotal13) + "\n";
            info += "Non-zero entries: " + std::to_string(st["non_zero_kmers"]) + "\n";
            info += "Total k-mer count: " + std::to_string(st["total_count"]) + "\n";
        } else if (ix23) {
            info += "Mode: 23-mer\n";
            info += "Total k-mers: " + std::to_string(checker.size()) + "\n";
        } else {
            info += "Mode: No index loaded\n";
        }
        if (aindex_loaded) {
            info += "AIndex: Loaded\n";
            info += "Reads: " + std::to_string(n_reads) + "\n";
        } else {
            info += "AIndex: Not loaded\n";
        }
        return info;
    }
    std::string get_23mer_statistics() {  // :1301-1315
        if (is_13mer_mode) return "Not in 23-mer mode";
        std::ostringstream s;
        s << "23-mer Index Statistics:\n";
        s << "Total k-mers: " << n_kmers << "\n";
        s << "Total reads: " << n_reads << "\n";
        s << "AIndex loaded: " << (aindex_loaded ? "Yes" : "No") << "\n";
        s << "Reads loaded: " << (reads != nullptr ? "Yes" : "No") << "\n";
        s << "Hash map size: " << checker.size() << "\n";
        return s.str();
    }
    void debug_kmer_tf_values() {  // :913-935
        for (uint64_t h : {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull}) {
            if (h >= checker.size()) continue;
            py::print(decode23(checker[h]), h, tf23[h]);
        }
    }

    // ------------------------------------------------------------------ builders (additions)
    // count_kmers13 <input> <pf> <out.tf.bin> (count_kmers13.cpp:546-612) -> stats dict
    py::dict count_kmers13(const std::string &input_file, const std::string &pf_file, const std::string &out_file) {
        require_file(input_file, "input file");
        require_file(pf_file, "hash file");
        ensure_ctx();
        Mapped in;
        in.open(input_file);
        aix_mphf *m = nullptr;
        locked([&] { return aix_mphf_load_pf(ctx, pf_file.c_str(), &m); });
        std::vector<uint64_t> tf(kTotal13);
        aix_count_stats st;
        int rc;
        {
            py::gil_scoped_release nogil;
            std::lock_guard<std::recursive_mutex> device_lock(mu);  // an aix_ctx serves one caller at a time (the reference relied on the GIL)
            rc = aix_count13(ctx, m, (const uint8_t *)in.ptr, in.size, AIX_FMT_DETECT, tf.data(), &st);
        }
        aix_mphf_destroy(ctx, m);
        check(rc);
        FILE *f = fopen(out_file.c_str(), "wb");
        if (!f) throw std::runtime_error("cannot create output file: " + out_file);
        bool ok = fwrite(tf.data(), 8, kTotal13, f) == kTotal13;
        ok = (fclose(f) == 0) && ok;
        if (!ok) throw std::runtime_error("short write: " + out_file);
        py::dict d;
        d["sequences"] = st.sequences;
        d["total_kmers"] = st.windows;
        d["valid_kmers"] = st.valid;
        d["invalid_kmers"] = st.invalid;
        return d;
    }

    // compute_aindex / compute_aindex13: positions index of the loaded index over a reads file,
    // written as {index_bin, indices_bin} (hash.hpp:470-486 / compute_aindex13.cpp:297-323)
    void build_positions(const std::string &reads_file, const std::string &index_bin, const std::string &indices_bin, int k = 23) {
        require_file(reads_file, "reads file");
        Mapped rd;
        rd.open(reads_file);
        uint64_t total = 0, n = 0;
        if (k == 23) {
            require23();
            locked([&] { return aix_positions_total23(ctx, ix23, &total); });
            n = checker.size();
        } else if (k == 13) {
            if (!ix13) throw std::runtime_error("13-mer index not loaded");
            locked([&] { return aix_positions_total13(ctx, ix13, &total); });
            n = kTotal13;
        } else {
            throw std::invalid_argument("k must be 13 or 23");
        }
        std::vector<uint64_t> indices(n + 1), positions(total);
        int rc;
        {
            py::gil_scoped_release nogil;
            std::lock_guard<std::recursive_mutex> device_lock(mu);  // an aix_ctx serves one caller at a time (the reference relied on the GIL)
            rc = k == 23 ? aix_positions_build23(ctx, ix23, (const uint8_t *)rd.ptr, rd.size, indices.data(), positions.data())
                         : aix_positions_build13(ctx, ix13, (const uint8_t *)rd.ptr, rd.size, indices.data(), positions.data());
        }
        check(rc);
        auto dump = [](const std::string &path, const std::vector<uint64_t> &v) {
            FILE *f = fopen(path.c_str(), "wb");
            if (!f) throw std::runtime_error("cannot create " + path);
            bool ok = fwrite(v.data(), 8, v.size(), f) == v.size();
            ok = (fclose(f) == 0) && ok;
            if (!ok) throw std::runtime_error("short write: " + path);
        };
        dump(index_bin, positions);
        dump(indices_bin, indices);
    }

    // reads file -> {prefix}.pf/.kmers.bin/.tf.bin entirely on the GPU (replaces the
    // jellyfish|kmer_counter -> compute_mphf_seq -> compute_index stages of scripts/compute_aindex.py)
    uint64_t build_index_from_reads(const std::string &reads_file, const std::string &prefix) {
        require_file(reads_file, "reads file");
        ensure_ctx();
        Mapped rd;
        rd.open(reads_file);
        uint64_t n = 0;
        locked([&] { return aix_canonical23_count(ctx, (const uint8_t *)rd.ptr, rd.size, &n, nullptr, nullptr); });
        std::vector<uint64_t> kmers(n), chk(n);
        std::vector<uint32_t> counts(n), tfv(n);
        locked([&] { return aix_canonical23_count(ctx, (const uint8_t *)rd.ptr, rd.size, &n, kmers.data(), counts.data()); });
        aix_mphf *m = nullptr;
        locked([&] { return aix_mphf_build(ctx, kmers.data(), n, 23, &m); });
        int rc;
        {
            std::lock_guard<std::recursive_mutex> device_lock(mu);
            rc = aix_index23_fill(ctx, m, kmers.data(), counts.data(), n, chk.data(), tfv.data());
            if (rc == AIX_OK) rc = aix_mphf_save_pf(ctx, m, (prefix + ".pf").c_str());
            aix_mphf_destroy(ctx, m);
        }
        check(rc);
        auto dump = [](const std::string &path, const void *p, size_t bytes) {
            FILE *f = fopen(path.c_str(), "wb");
            if (!f) throw std::runtime_error("cannot create " + path);
            bool ok = fwrite(p, 1, bytes, f) == bytes;
            ok = (fclose(f) == 0) && ok;
            if (!ok) throw std::runtime_error("short write: " + path);
        };
        dump(prefix + ".kmers.bin", chk.data(), n * 8);
        dump(prefix + ".tf.bin", tfv.data(), n * 4);
        return n;
    }
};

PYBIND11_MODULE(aindex_cpp, m) {
    m.doc() = "aindex_cpp: the AindexWrapper API of ad3002/aindex on libaindex_cuda (B200, sm_100a)";
    py::register_exception<FileNotFound>(m, "AindexFileNotFound", PyExc_FileNotFoundError);
    m.attr("backend") = "cuda-sm_100a";
    m.attr("version") = aix_version();

    py::class_<AindexWrapper>(m, "AindexWrapper")
        .def(py::init<>())
        .def("load", &AindexWrapper::load)
        .def("load_hash_file", &AindexWrapper::load_hash_file)
        .def("load_reads", &AindexWrapper::load_reads)
        .def("load_reads_index", &AindexWrapper::load_reads_index)
        .def("load_reads_in_memory", &AindexWrapper::load_reads_in_memory)
        .def("load_aindex", &AindexWrapper::load_aindex)
        .def("load_13mer_index", &AindexWrapper::load_13mer_index)
        .def("load_13mer_aindex", &AindexWrapper::load_13mer_aindex)
        .def("load_from_prefix_23mer", &AindexWrapper::load_from_prefix_23mer, py::arg("prefix"), py::arg("reads_file") = "")
        .def("load_aindex_from_prefix_23mer", &AindexWrapper::load_aindex_from_prefix_23mer, py::arg("prefix"), py::arg("max_tf"),
             py::arg("reads_file") = "")
        .def("load_from_prefix_13mer", &AindexWrapper::load_from_prefix_13mer, py::arg("prefix"), py::arg("reads_file") = "")
        .def("load_aindex_from_prefix_13mer", &AindexWrapper::load_aindex_from_prefix_13mer, py::arg("prefix"),
             py::arg("reads_file") = "")
        .def("get_tf_values", &AindexWrapper::tf_values_list, "list[str] -> list[int] (python_wrapper.cpp:653-664)")
        .def("get_tf_values", &AindexWrapper::get_tf_values)
        .def("get_tf_values", &AindexWrapper::get_tf_values_array, "uint8[q, k] records -> uint32[q] (no per-string objects)")
        .def("get_tf_value", &AindexWrapper::get_tf_value)
        .def("get_hash_values", &AindexWrapper::get_hash_values)
        .def("get_hash_value", &AindexWrapper::get_hash_value)
        .def("get_kid_by_kmer", &AindexWrapper::get_kid_by_kmer)
        .def("get_kmer_by_kid", &AindexWrapper::get_kmer_by_kid)
        .def("get_strand", &AindexWrapper::get_strand)
        .def("get_kmer_info", &AindexWrapper::get_kmer_info)
        .def("get_rid", &AindexWrapper::get_rid)
        .def("get_start", &AindexWrapper::get_start)
        .def("get_read_by_rid", &AindexWrapper::get_read_by_rid)
        .def("get_reads_by_rids", &AindexWrapper::get_reads_by_rids, py::arg("rids"),
             "batched get_read_by_rid: (concatenated bytes, uint64 offsets[q+1])")
        .def("get_rids_and_starts", &AindexWrapper::get_rids_and_starts, py::arg("positions"),
             "batched get_rid / get_start: (uint64 rids[q], uint64 starts[q])")
        .def("get_read", &AindexWrapper::get_read, py::arg("start"), py::arg("end"), py::arg("revcomp") = false)
        .def("get_reads_se_by_kmer", &AindexWrapper::get_reads_se_by_kmer)
        .def("get_positions", &AindexWrapper::get_positions)
        .def("get_positions_13mer", &AindexWrapper::get_positions_13mer)
        .def("get_positions_batch", &AindexWrapper::get_positions_batch, py::arg("kmers"), py::arg("k") = 23)
        .def("get_hash_size", &AindexWrapper::get_hash_size)
        .def("get_reads_size", &AindexWrapper::get_reads_size)
        .def_readwrite("aindex_loaded", &AindexWrapper::aindex_loaded)
        .def_readwrite("n_reads", &AindexWrapper::n_reads)
        .def_readwrite("n_kmers", &AindexWrapper::n_kmers)
        .def_readwrite("reads_size", &AindexWrapper::reads_size)
        .def("debug_kmer_tf_values", &AindexWrapper::debug_kmer_tf_values)
        .def("get_index_info", &AindexWrapper::get_index_info)
        .def("get_total_tf_value_13mer", &AindexWrapper::get_total_tf_value_13mer)
        .def("get_total_tf_values_13mer", &AindexWrapper::total_tf_values_13mer_list)
        .def("get_total_tf_values_13mer", &AindexWrapper::get_total_tf_values_13mer)
        .def("get_tf_both_directions_13mer", &AindexWrapper::get_tf_both_directions_13mer)
        .def("get_tf_both_directions_13mer_batch", &AindexWrapper::get_tf_both_directions_13mer_batch)
        .def("get_reverse_complement_13mer", &AindexWrapper::get_reverse_complement_13mer)
        .def("get_13mer_statistics", &AindexWrapper::get_13mer_statistics)
        .def("get_13mer_tf_array", &AindexWrapper::get_13mer_tf_array)
        .def("get_13mer_tf_array_direct", &AindexWrapper::get_13mer_tf_array_direct, "uint64[4^13]: tf by 2-bit 13-mer value")
        .def("get_tf_array_23mer", &AindexWrapper::get_tf_array_23mer, "uint32[n]: tf by kid")
        .def("get_tf_by_index_13mer", &AindexWrapper::get_tf_by_index_13mer)
        .def("get_tf_values_13mer", &AindexWrapper::tf_values_13mer_list)
        .def("get_tf_values_13mer", &AindexWrapper::get_tf_values_13mer)
        .def("get_tf_values_23mer", &AindexWrapper::tf_values_23mer_list)
        .def("get_tf_values_23mer", &AindexWrapper::get_tf_values_23mer)
        .def("get_total_tf_value_23mer", &AindexWrapper::get_total_tf_value_23mer)
        .def("get_total_tf_values_23mer", &AindexWrapper::total_tf_values_23mer_list)
        .def("get_total_tf_values_23mer", &AindexWrapper::get_total_tf_values_23mer)
        .def("get_tf_both_directions_23mer", &AindexWrapper::get_tf_both_directions_23mer)
        .def("get_tf_both_directions_23mer_batch", &AindexWrapper::get_tf_both_directions_23mer_batch)
        .def("get_reverse_complement_23mer", &AindexWrapper::get_reverse_complement_23mer)
        .def("get_23mer_statistics", &AindexWrapper::get_23mer_statistics)
        // additions
        .def("get_sequence_coverage", &AindexWrapper::get_sequence_coverage, py::arg("seq"), py::arg("cutoff") = 0, py::arg("k") = 23)
        .def("get_sequence_coverage_batch", &AindexWrapper::get_sequence_coverage_batch, py::arg("seqs"), py::arg("offsets"),
             py::arg("cutoff") = 0, py::arg("k") = 23)
        .def("count_kmers13", &AindexWrapper::count_kmers13, py::arg("input_file"), py::arg("pf_file"), py::arg("out_file"))
        .def("build_positions", &AindexWrapper::build_positions, py::arg("reads_file"), py::arg("index_bin"), py::arg("indices_bin"),
             py::arg("k") = 23)
        .def("build_index_from_reads", &AindexWrapper::build_index_from_reads, py::arg("reads_file"), py::arg("prefix"));
}
