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

// whole file into memory (c2b_fastq.cpp): plain read / gzip inflate
bool c2b_fastq_read_gz(const char *path, std::vector<uint8_t> &buf, std::string &err);
void c2b_fastq_set_error(const std::string &m);
