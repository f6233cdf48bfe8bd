// c2b_fastq_int.h -- result object shared by the host (c2b_fastq.cpp) and GPU (c2b_fastq_gpu.cu) FASTQ front ends.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

struct c2b_fastq {
    std::unique_ptr<uint8_t[]> seqs;                        // packed unique sequences (uninitialised storage: every byte is written by emit)
    std::vector<int64_t> offsets;
    std::vector<int32_t> counts;
    std::vector<int64_t> first_index;
    int64_t n_reads = 0;
    int32_t max_len = 0;
    std::string err;
};

// byte buffer whose resize() leaves new bytes uninitialised (a whole FASTQ is read or inflated into it: no memset pass first)
template <class T> struct c2b_noinit_alloc : std::allocator<T> {
    template <class U> struct rebind { using other = c2b_noinit_alloc<U>; };
    template <class U> void construct(U *p) { ::new ((void *)p) U; }
    template <class U, class A, class... As> void construct(U *p, A &&a, As &&...as) { ::new ((void *)p) U(std::forward<A>(a), std::forward<As>(as)...); }
};
using c2b_bytes = std::vector<uint8_t, c2b_noinit_alloc<uint8_t>>;

// whole file into memory (c2b_fastq.cpp): plain read / gzip inflate (blocked gzip: all host threads)
bool c2b_fastq_read_gz(const char *path, c2b_bytes &buf, std::string &err);
void c2b_fastq_set_error(const std::string &m);
