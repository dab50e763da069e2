// compute_reads <fastq_file1|fasta_file1|reads_file> <fastq_file2|-> <fastq|fasta|se|reads> <output_prefix>
// Same positional arguments and byte-identical outputs as the reference tool (src/compute_reads.cpp:20-224):
//   {prefix}.reads   one read per line; paired reads as read1 ~ revcomp(read2) (:85-96, get_revcomp kmers.cpp:309-329)
//   {prefix}.ridx    "rid\tstart\tend" per read (:99, :121, :145, :176, :196)
//   {prefix}.header  FASTA only: "header\tstart\tlength" (:178, :198)
// Sequential text conversion: there is nothing data-parallel to put on the GPU here, so this is host C++ working on the
// memory-mapped inputs with one buffered writer per output; it exists so that the whole index pipeline (compute_reads ->
// count -> index -> positions) can run from this package's bin/ directory.
#include <errno.h>

#include "tool_common.hpp"

struct LineReader {  // std::getline over a mapped file (a final line without '\n' counts; '\r' is kept, as getline does)
    const uint8_t *p, *e;
    explicit LineReader(const MappedFile &f) : p(f.data), e(f.data + f.size) {}
    bool next(const uint8_t *&b, size_t &n) {
        if (p == nullptr || p >= e) return false;
        const uint8_t *nl = (const uint8_t *)memchr(p, '\n', (size_t)(e - p));
        b = p;
        n = nl ? (size_t)(nl - p) : (size_t)(e - p);
        p = nl ? nl + 1 : e;
        return true;
    }
};

struct Writer {
    FILE *f = nullptr;
    std::vector<char> buf;
    bool open(const std::string &path) {
        f = fopen(path.c_str(), "wb");
        if (!f) return false;
        buf.resize(8 << 20);
        setvbuf(f, buf.data(), _IOFBF, buf.size());
        return true;
    }
    void put(const void *p, size_t n) { if (n) fwrite(p, 1, n, f); }
    void put(char c) { fputc(c, f); }
    void num(unsigned long long v) { fprintf(f, "%llu", v); }
    bool close() { bool ok = f && !ferror(f); if (f) ok = (fclose(f) == 0) && ok; f = nullptr; return ok; }
};

static void make_dirs(const std::string &prefix) {  // :36-63
    size_t last = prefix.find_last_of('/');
    if (last == std::string::npos) return;
    std::string dir = prefix.substr(0, last), acc;
    size_t i = 0;
    if (!dir.empty() && dir[0] == '/') acc = "";
    while (i <= dir.size()) {
        size_t j = dir.find('/', i);
        if (j == std::string::npos) j = dir.size();
        std::string seg = dir.substr(i, j - i);
        if (!seg.empty()) {
            acc += (acc.empty() && dir[0] != '/') ? seg : "/" + seg;
            struct stat st;
            if (stat(acc.c_str(), &st) != 0 && mkdir(acc.c_str(), 0755) != 0 && errno != EEXIST) {
                fprintf(stderr, "Error creating directory %s: %s\n", acc.c_str(), strerror(errno));
                exit(1);
            }
        }
        i = j + 1;
    }
}

static void ridx_line(Writer &w, uint64_t rid, uint64_t s, uint64_t e) {
    w.num(rid); w.put('\t'); w.num(s); w.put('\t'); w.num(e); w.put('\n');
}

int main(int argc, char **argv) {
    if (argc < 5) {
        fprintf(stderr, "Convert fasta or fastq reads to simple reads.\nExpected arguments: %s <fastq_file1|fasta_file1|reads_file> "
                        "<fastq_file2|-> <fastq|fasta|se|reads> <output_prefix>\n", argv[0]);
        return 1;
    }
    const std::string type = argv[3], prefix = argv[4];
    make_dirs(prefix);
    MappedFile f1, f2;
    if (!f1.open(argv[1])) { fprintf(stderr, "Error: Cannot open input file: %s\n", argv[1]); return 1; }
    Writer reads, ridx, header;
    uint64_t n_reads = 0, start = 0;
    const uint8_t *a, *b;
    size_t na, nb;
    if (type == "fastq") {
        if (!f2.open(argv[2])) { fprintf(stderr, "Error: Cannot open input file: %s\n", argv[2]); return 1; }
        if (!reads.open(prefix + ".reads") || !ridx.open(prefix + ".ridx")) return 1;
        LineReader r1(f1), r2(f2);
        std::vector<char> rc;
        while (r1.next(a, na)) {                       // header of read 1
            const bool have1 = r1.next(a, na);         // sequence 1 (an absent line reads as empty, as a failed getline leaves "")
            if (!have1) na = 0;
            r2.next(b, nb);                            // header of read 2
            if (!r2.next(b, nb)) nb = 0;               // sequence 2
            const uint64_t end = start + na + nb + 1;  // + '~'
            rc.resize(nb);
            for (size_t y = 0; y < nb; ++y) {          // get_revcomp(const std::string&): ACGT complemented, anything else N
                const uint8_t c = b[nb - 1 - y];
                rc[y] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N';
            }
            reads.put(a, na); reads.put('~'); reads.put(rc.data(), nb); reads.put('\n');
            ridx_line(ridx, n_reads, start, end);
            start = end + 1;
            r1.next(a, na); r1.next(a, na); r2.next(b, nb); r2.next(b, nb);  // '+' and quality lines
            ++n_reads;
        }
    } else if (type == "se") {
        if (!reads.open(prefix + ".reads") || !ridx.open(prefix + ".ridx")) return 1;
        LineReader r1(f1);
        while (r1.next(a, na)) {
            if (!r1.next(a, na)) na = 0;
            const uint64_t end = start + na;
            reads.put(a, na); reads.put('\n');
            ridx_line(ridx, n_reads, start, end);
            start = end + 1;
            r1.next(a, na); r1.next(a, na);
            ++n_reads;
        }
    } else if (type == "reads") {
        if (!ridx.open(prefix + ".ridx")) return 1;
        LineReader r1(f1);
        while (r1.next(a, na)) {
            ridx_line(ridx, n_reads, start, start + na);
            start += na + 1;
            ++n_reads;
        }
    } else if (type == "fasta") {
        if (!reads.open(prefix + ".reads") || !ridx.open(prefix + ".ridx") || !header.open(prefix + ".header")) return 1;
        LineReader r1(f1);
        std::string seq, head;
        auto flush = [&]() {
            const uint64_t end = start + seq.size();
            reads.put(seq.data(), seq.size()); reads.put('\n');
            ridx_line(ridx, n_reads, start, end);
            header.put(head.data(), head.size()); header.put('\t'); header.num(start); header.put('\t'); header.num(seq.size()); header.put('\n');
            start = end + 1;
            ++n_reads;
            seq.clear();
        };
        while (r1.next(a, na)) {
            if (na && a[0] == '>') {       // (an empty line is line1[0] == '\0' in the reference: not a header)
                if (!seq.empty()) flush();
                head.assign((const char *)a + 1, na - 1);
                continue;
            }
            seq.append((const char *)a, na);
        }
        if (!seq.empty()) flush();
    } else {
        fprintf(stderr, "Unknown format.\n");
        return 2;
    }
    bool ok = true;
    if (reads.f) ok = reads.close() && ok;
    if (ridx.f) ok = ridx.close() && ok;
    if (header.f) ok = header.close() && ok;
    return ok ? 0 : 1;
}
