// usb_hostindex.h -- host-side construction of the UDB index (CSR postings) and masked SeqDB.
//
// Replaces, for the in-memory layout: MaskDB (makeudb.cpp:11-25) -> FastMaskSeq
// (fastmask.cpp:88-158), UDBParams::SetTargetWords/SetTargetUniqueWords (udbparams.cpp:644-711)
// and UDBData::FromSeqDB (udbbuild.cpp:303-398).  The reference keeps one heap row per word
// (m_UDBRows[word], m_Sizes[word]); here the rows are one CSR array so that a posting row is a
// contiguous, coalesced HBM range.
#pragma once
#include <stdint.h>
#include <vector>
#include "usb_tables.h"

namespace usb {

struct HostIndex {
	uint32_t n_seq = 0;
	uint32_t word_length = 8;
	uint32_t slots = 0;                 // 4^word_length
	std::vector<uint8_t> seqs;          // masked letters, each target padded to a 16-byte boundary
	std::vector<uint64_t> seq_off;      // n_seq+1 padded offsets (multiples of 16)
	std::vector<uint32_t> seq_len;      // n_seq true lengths
	std::vector<uint64_t> row_off;      // slots+1; every row starts on a 16-byte boundary (multiple of 4 entries)
	std::vector<uint32_t> row_size;     // slots; true row lengths (m_Sizes[word])
	std::vector<uint32_t> postings;     // target indexes, ascending per row, each target once per row
	uint64_t n_postings = 0;            // sum of row_size
	uint32_t max_len = 0;
};

// FastMaskSeq soft-masking (fastmask.cpp:88-158); in == out allowed.
void fastmask_nt(const uint8_t *in, uint32_t L, uint8_t *out);

// Builds the masked SeqDB + CSR index.  dbmask: 1 = fastnucleo, 0 = sequences taken verbatim
// (cluster_fast indexes raw reads, clusterfast.cpp:88-103).  n_threads <= 0: hardware concurrency.
void build_host_index(const uint8_t *seqs, const uint64_t *seq_off, uint32_t n_seq, uint32_t word_length,
  int dbmask, int n_threads, HostIndex &out);

} // namespace usb
