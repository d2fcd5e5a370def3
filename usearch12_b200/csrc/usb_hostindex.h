// usb_hostindex.h -- host-side construction of the UDB index (CSR postings) and masked SeqDB.
//
// Replaces, for the in-memory layout: MaskDB (makeudb.cpp:11-25) -> FastMaskSeq
// (fastmask.cpp:88-158), UDBParams::SetTargetWords/SetTargetUniqueWords (udbparams.cpp:644-711),
// UDBData::FromSeqDB (udbbuild.cpp:303-398) and, for the growing cluster_fast database,
// UDBData::AddSIToDB_CopyData / AddWord / GrowRow (udbbuild.cpp:74-130,286).
// The reference keeps one heap row per word (m_UDBRows[word], m_Sizes[word]) and grows rows in
// place; here the index is a short list of immutable CSR segments over consecutive target ranges
// (a log-structured merge: appended targets form a new small segment, neighbouring segments of
// similar size are concatenated), so that every posting row fragment stays a contiguous, 16-byte
// aligned HBM range and an append never rewrites more than an amortised O(log N) share.
#pragma once
#include <stdint.h>
#include <vector>
#include "usb_tables.h"

namespace usb {

// Masked target letters; each target padded to a 16-byte boundary.
struct HostSeqs {
	std::vector<uint8_t> seqs;
	std::vector<uint64_t> seq_off{0};   // n+1 padded offsets (multiples of 16)
	std::vector<uint32_t> seq_len;      // true lengths
	uint32_t max_len = 0;
	uint32_t n() const { return (uint32_t)seq_len.size(); }
	// dbmask: 1 = fastnucleo soft-masking, 0 = letters taken verbatim (cluster_fast)
	void append(const uint8_t *s, const uint64_t *off, uint32_t count, int dbmask, int n_threads);
};

// Postings of the targets [base, base+count): row_off[slots+1] (multiples of 4 entries),
// row_size[slots] true lengths, postings = global target indexes ascending per row, every target
// at most once per row; padding entries are 0xffffffff.
struct HostCSR {
	uint32_t base = 0, count = 0, slots = 0;
	std::vector<uint64_t> row_off;
	std::vector<uint32_t> row_size;
	std::vector<uint32_t> postings;
	uint64_t n_postings = 0;
};

// The same index as 2-byte "increment descriptors" laid out for the way a warp of k_rank walks a
// row, for static indexes of at most USB_HALF_MAX_TARGETS targets starting at target 0.  It halves
// the HBM bytes of the U-sort walk, removes most shared-memory bank conflicts of the counter
// increments and cuts the walk to three instructions per posting (extract, address, ATOMS),
// which is what bounds it on B200 (profiles/README.md).
//   * The counter of target t is byte t % 4 of the 32-bit shared-memory word t / 4.  An entry is
//     the word index t / 4 (16 bits); the byte is implied by the entry's position.
//   * A row is a whole number of groups of 256 entries = 32 lanes x one 128-bit vector.  Lane l
//     loads vector 32 g + l of group g; the warp then issues one increment instruction per entry
//     index i = 0..7, i.e. for the entries {256 g + 8 l + i : l}.  Entries i = 2b and 2b + 1 hold
//     targets with t % 4 == b, so the increment value 1 << 8b is a constant of the instruction.
//   * The order of the postings inside a row does not matter for U: the targets of each byte
//     class are dealt round-robin over the 32 banks (bank = (t / 4) % 32), so the 32 increments
//     of one instruction fall into (mostly) different banks.
//   * Unused entries of lane l point at dummy word dummy0 + l behind the counters (one per bank),
//     so the walk has no validity tests at all.  row_groups[w] = max over the byte classes of
//     ceil(n_class / 64): about 7 % padding at 2 146 postings per row.
#define USB_HALF_MAX_TARGETS 200000u
struct HostHalf {
	uint32_t dummy0 = 0;             // first dummy word = round16(n_targets) / 4
	std::vector<uint64_t> row_off;   // slots + 1, in entries (multiples of 256)
	std::vector<uint32_t> row_groups; // slots
	std::vector<uint16_t> postings;
};
void make_half(const HostCSR &H, uint32_t n_targets, int n_threads, HostHalf &out);

// FastMaskSeq soft-masking (fastmask.cpp:88-158); in == out allowed.
void fastmask_nt(const uint8_t *in, uint32_t L, uint8_t *out);

// UDB alphabet (udbparams.cpp:235-261): alpha = 4 (ACGT, U = T) or 20 (ACDEFGHIKLMNPQRSTVWY);
// a word is `word_length` letters read as a base-alpha number, slots = alpha^word_length.
// Returns the 256-entry character -> letter table (0xff = lower case / not in the alphabet).
const uint8_t *udb_letters(uint32_t alpha);
uint32_t udb_slots(uint32_t alpha, uint32_t word_length);
// word starting at s, or UINT32_MAX when a letter is bad (udbparams.cpp:540-555)
inline uint32_t udb_word(const uint8_t *s, uint32_t w, uint32_t alpha, const uint8_t *letters)
{
	uint32_t word = 0;
	for (uint32_t i = 0; i < w; ++i) {
		const uint32_t l = letters[s[i]];
		if (l == 0xff)
			return UINT32_MAX;
		word = word * alpha + l;
	}
	return word;
}

// n_threads <= 0: hardware concurrency.
void build_csr(const HostSeqs &S, uint32_t first, uint32_t count, uint32_t word_length, uint32_t alpha, int n_threads,
  HostCSR &out);
// a covers [x, y), b covers [y, z): out covers [x, z) with rows = a's row followed by b's row.
void merge_csr(const HostCSR &a, const HostCSR &b, HostCSR &out);

} // namespace usb
