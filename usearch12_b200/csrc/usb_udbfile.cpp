// usb_udbfile.cpp -- .udb database files of the reference, host only (no device needed).
//
// File layout reproduced byte for byte (all little endian, structs packed as the reference's):
//   UDBFileHdr                      udbfile.h:17-62, filled by UDBFileHdr::FromParams (udbio.cpp:13-57)
//   uint32 sizes[slots]             UDBData::ToUDBFile                        udbio.cpp:283-318
//   uint32 'UDB3'
//   rows: sizes[w] uint32 targets   WriteRowsNotVarCoded (ascending target indexes)
//   uint32 'UDB4'
//   SeqDB::ToFile                   seqdbio.cpp:17-135: SeqDBFileHdr (32 bytes with tail padding),
//                                   label offsets, labels (NUL terminated), lengths, letters
// usb_udb_write is -makeudb_usearch (makeudb.cpp:27-60): MaskDB -> UDBData::FromSeqDB -> ToUDBFile
// for the default, non-hashed, non-coded index (word width 8 for nt, 5 for aa; -dbstep 1,
// -dbaccelpct 100).  The stored sequences are the masked ones (lower case = masked), which is why
// the reference does not mask again when it loads a .udb (loaddb.cpp:100-127).
#include <cstdio>
#include <cstring>
#include <string>
#include <memory>
#include <new>
#include <vector>

#include "../../include/usb200.h"
#include "usb_hostindex.h"

#include <cstdarg>

namespace usb {
int fail_msg(int code, const char *msg); // usb_api.cu: sets usb_last_error()
}
using namespace usb;

static int fail(int code, const char *fmt, ...)
{
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	return fail_msg(code, buf);
}

namespace {
#pragma pack(push, 1)
struct UdbHdr {
	uint32_t magic1, hashed, seq_index_bits, seq_pos_bits, word_width, db_step, db_accel_pct, rfu1, rfu2, utax_data, end_of_row;
	uint64_t slot_count, seq_count;
	uint8_t step_prefix[8];
	char alpha_str[64], pattern_str[64];
	uint32_t magic2;
};
#pragma pack(pop)
struct SeqDbHdr { // natural alignment, 32 bytes (seqdb.h:19-27)
	uint32_t magic1, seq_count;
	uint64_t seq_bytes;
	uint32_t label_bytes, split_count, magic2, pad;
};
static_assert(sizeof(UdbHdr) == 200, "UDBFileHdr is 200 bytes");
static_assert(sizeof(SeqDbHdr) == 32, "SeqDBFileHdr is 32 bytes");
#define MAGIC4(a, b, c, d) ((uint32_t)(a) << 24 | (uint32_t)(b) << 16 | (uint32_t)(c) << 8 | (uint32_t)(d))
const uint32_t UDB_MAGIC1 = MAGIC4('U', 'D', 'B', 'F'), UDB_MAGIC2 = MAGIC4('U', 'D', 'B', 'f');
const uint32_t UDB_MAGIC3 = MAGIC4('U', 'D', 'B', '3'), UDB_MAGIC4 = MAGIC4('U', 'D', 'B', '4');
const uint32_t SEQDB_MAGIC1 = 0x5E0DB3, SEQDB_MAGIC2 = 0x5E0DB4;

struct File {
	FILE *f = nullptr;
	~File()
	{
		if (f)
			fclose(f);
	}
};
} // namespace

struct usb_udb {
	bool nucleo = true;
	uint32_t word_length = 8, slots = 0;
	std::vector<uint32_t> sizes;
	std::vector<uint64_t> row_off; // slots + 1
	std::vector<uint32_t> rows;
	std::vector<uint8_t> seqs;
	std::vector<uint64_t> seq_off{0};
	std::vector<char> labels;
	std::vector<uint32_t> label_off;
};

extern "C" int usb_udb_write(const char *path, const usb_params *p, const uint8_t *seqs, const uint64_t *seq_off,
  const char *const *labels, uint32_t n_seq)
{
	if (!path || !p || !seq_off || !labels || (!seqs && n_seq))
		return fail(USB_EINVAL, "usb_udb_write: null argument");
	if (p->struct_size != sizeof(usb_params))
		return fail(USB_EINVAL, "usb_params.struct_size mismatch (header/library version skew)");
	if (n_seq == 0)
		return fail(USB_EINVAL, "Empty database"); // udbio.cpp:340-341
	const uint32_t alpha = p->is_nucleo ? 4 : 20;
	const uint32_t slots = udb_slots(alpha, p->word_length);
	if (slots == 0 || (p->is_nucleo ? p->word_length > 8 : p->word_length > 5))
		return fail(USB_EINVAL, "word_length %u unsupported", p->word_length);
	HostSeqs S;
	S.append(seqs, seq_off, n_seq, p->dbmask, 0);
	HostCSR H;
	build_csr(S, 0, n_seq, p->word_length, alpha, 0, H);
	File out;
	out.f = fopen(path, "wb");
	if (!out.f)
		return fail(USB_EINVAL, "Cannot create %s", path);
	bool ok = true;
	auto put = [&](const void *d, size_t n) { ok = ok && (n == 0 || fwrite(d, 1, n, out.f) == n); };
	UdbHdr h;
	memset(&h, 0, sizeof h);
	h.magic1 = UDB_MAGIC1;
	h.seq_index_bits = 32;
	h.word_width = p->word_length;
	h.db_step = 1;
	h.db_accel_pct = 100;
	h.seq_count = n_seq;
	strcpy(h.alpha_str, p->is_nucleo ? "nt" : "aa");
	h.magic2 = UDB_MAGIC2;
	put(&h, sizeof h);
	put(H.row_size.data(), (size_t)slots * 4);
	put(&UDB_MAGIC3, 4);
	for (uint32_t w = 0; w < slots; ++w)
		put(H.postings.data() + H.row_off[w], (size_t)H.row_size[w] * 4);
	put(&UDB_MAGIC4, 4);
	// SeqDB::ToFile
	SeqDbHdr sh;
	memset(&sh, 0, sizeof sh);
	sh.magic1 = SEQDB_MAGIC1;
	sh.seq_count = n_seq;
	sh.magic2 = SEQDB_MAGIC2;
	std::vector<uint32_t> loff(n_seq);
	std::vector<char> lbuf;
	for (uint32_t i = 0; i < n_seq; ++i) {
		loff[i] = (uint32_t)lbuf.size();
		const char *l = labels[i] ? labels[i] : "";
		lbuf.insert(lbuf.end(), l, l + strlen(l) + 1);
		sh.seq_bytes += S.seq_len[i];
	}
	if (lbuf.size() > 0xfffffbffull)
		return fail(USB_ELIMIT, "Label data too big"); // seqdbio.cpp:12-13
	sh.label_bytes = (uint32_t)lbuf.size();
	put(&sh, sizeof sh);
	put(loff.data(), (size_t)n_seq * 4);
	put(lbuf.data(), lbuf.size());
	put(S.seq_len.data(), (size_t)n_seq * 4);
	for (uint32_t i = 0; i < n_seq; ++i)
		put(S.seqs.data() + S.seq_off[i], S.seq_len[i]);
	if (!ok)
		return fail(USB_EINVAL, "Write error on %s", path);
	return 0;
}

extern "C" int usb_udb_probe(const char *path)
{
	File in;
	in.f = path ? fopen(path, "rb") : nullptr;
	if (!in.f)
		return 0;
	uint32_t m = 0;
	return fread(&m, 4, 1, in.f) == 1 && m == UDB_MAGIC1;
}

// The sizes in a .udb header are untrusted: every allocation is checked against the bytes left in
// the file first, and allocation failures are reported as USB_ENOMEM (no exception crosses the ABI).
static int udb_read_checked(const char *path, usb_udb **out)
{
	File in;
	in.f = fopen(path, "rb");
	if (!in.f)
		return fail(USB_EINVAL, "Cannot open %s", path);
	uint64_t file_bytes = 0;
	if (fseek(in.f, 0, SEEK_END) == 0) {
		const long e = ftell(in.f);
		file_bytes = e > 0 ? (uint64_t)e : 0;
	}
	rewind(in.f);
	bool ok = true;
	uint64_t consumed = 0;
	auto get = [&](void *d, size_t n) {
		ok = ok && (n == 0 || fread(d, 1, n, in.f) == n);
		consumed += n;
	};
	auto left = [&]() { return file_bytes > consumed ? file_bytes - consumed : 0; };
	UdbHdr h;
	get(&h, sizeof h);
	if (!ok || h.magic1 != UDB_MAGIC1 || h.magic2 != UDB_MAGIC2)
		return fail(USB_EINVAL, "%s is not a .udb file (bad header magic)", path);
	// UDBFileHdr::ValidateFeatures: only the plain usearch index is understood here
	if (h.hashed || h.seq_pos_bits != 0 || h.pattern_str[0] || h.db_step != 1 || h.end_of_row || h.utax_data ||
	    h.db_accel_pct != 100)
		return fail(USB_EINVAL, "%s: hashed / coded / spaced / stepped / utax .udb variants are not supported", path);
	const bool nucleo = strcmp(h.alpha_str, "nt") == 0;
	if (!nucleo && strcmp(h.alpha_str, "aa") != 0)
		return fail(USB_EINVAL, "%s: alphabet '%.60s' is not supported", path, h.alpha_str);
	const uint32_t slots = udb_slots(nucleo ? 4 : 20, h.word_width);
	if (slots == 0 || (nucleo ? h.word_width > 8 : h.word_width > 5) || h.seq_count > 0xfffffff0ull)
		return fail(USB_EINVAL, "%s: word width %u / %llu sequences unsupported", path, h.word_width, (unsigned long long)h.seq_count);
	if ((uint64_t)slots * 4 > left())
		return fail(USB_EINVAL, "%s: truncated or inconsistent .udb file", path);
	std::unique_ptr<usb_udb> u(new usb_udb);
	u->nucleo = nucleo;
	u->word_length = h.word_width;
	u->slots = slots;
	u->sizes.resize(slots);
	get(u->sizes.data(), (size_t)slots * 4);
	uint32_t m = 0;
	get(&m, 4);
	if (!ok || m != UDB_MAGIC3)
		return fail(USB_EINVAL, "%s: .udb magic3 %08x should be %08x", path, m, UDB_MAGIC3);
	u->row_off.assign((size_t)slots + 1, 0);
	for (uint32_t w = 0; w < slots; ++w)
		u->row_off[w + 1] = u->row_off[w] + u->sizes[w];
	if (u->row_off[slots] > left() / 4)
		return fail(USB_EINVAL, "%s: truncated or inconsistent .udb file (rows of %llu entries)", path,
		  (unsigned long long)u->row_off[slots]);
	u->rows.resize(u->row_off[slots]);
	get(u->rows.data(), u->rows.size() * 4);
	get(&m, 4);
	if (!ok || m != UDB_MAGIC4)
		return fail(USB_EINVAL, "%s: .udb magic4 0x%08x should be 0x%08x", path, m, UDB_MAGIC4);
	SeqDbHdr sh;
	get(&sh, sizeof sh);
	if (!ok || sh.magic1 != SEQDB_MAGIC1 || sh.magic2 != SEQDB_MAGIC2 || sh.seq_count != h.seq_count)
		return fail(USB_EINVAL, "%s: SeqDB::FromFile, invalid header magics %08x %08x", path, sh.magic1, sh.magic2);
	const uint32_t n = sh.seq_count;
	if ((uint64_t)n * 8 > left() || (uint64_t)sh.label_bytes > left() - (uint64_t)n * 8 ||
	    (uint64_t)sh.seq_bytes > left() - (uint64_t)n * 8 - sh.label_bytes)
		return fail(USB_EINVAL, "%s: truncated or inconsistent .udb file", path);
	u->label_off.resize(n);
	get(u->label_off.data(), (size_t)n * 4);
	u->labels.resize((size_t)sh.label_bytes + 1, 0);
	get(u->labels.data(), sh.label_bytes);
	std::vector<uint32_t> len(n);
	get(len.data(), (size_t)n * 4);
	u->seq_off.assign((size_t)n + 1, 0);
	for (uint32_t i = 0; i < n; ++i)
		u->seq_off[i + 1] = u->seq_off[i] + len[i];
	if (ok && u->seq_off[n] != sh.seq_bytes)
		ok = false;
	if (ok) {
		u->seqs.resize(sh.seq_bytes + 1);
		get(u->seqs.data(), sh.seq_bytes);
	}
	for (uint32_t i = 0; ok && i < n; ++i)
		ok = u->label_off[i] < sh.label_bytes;
	if (!ok)
		return fail(USB_EINVAL, "%s: truncated or inconsistent .udb file", path);
	*out = u.release();
	return 0;
}

extern "C" int usb_udb_read(const char *path, usb_udb **out)
{
	if (!path || !out)
		return fail(USB_EINVAL, "usb_udb_read: null argument");
	try {
		return udb_read_checked(path, out);
	} catch (const std::bad_alloc &) {
		return fail(USB_ENOMEM, "%s: out of memory while reading the .udb file", path);
	}
}

extern "C" void usb_udb_free(usb_udb *u) { delete u; }
extern "C" uint32_t usb_udb_seq_count(const usb_udb *u) { return u ? (uint32_t)u->seq_off.size() - 1 : 0; }
extern "C" int usb_udb_is_nucleo(const usb_udb *u) { return u && u->nucleo; }
extern "C" uint32_t usb_udb_word_length(const usb_udb *u) { return u ? u->word_length : 0; }
extern "C" const uint8_t *usb_udb_seqs(const usb_udb *u, const uint64_t **seq_off)
{
	if (!u)
		return nullptr;
	if (seq_off)
		*seq_off = u->seq_off.data();
	return u->seqs.data();
}
extern "C" const char *usb_udb_label(const usb_udb *u, uint32_t i)
{
	return u && i < u->label_off.size() ? u->labels.data() + u->label_off[i] : nullptr;
}
extern "C" int usb_udb_row(const usb_udb *u, uint32_t word, const uint32_t **row, uint32_t *size)
{
	if (!u || word >= u->slots || !row || !size)
		return fail(USB_EINVAL, "usb_udb_row: bad argument");
	*row = u->rows.data() + u->row_off[word];
	*size = u->sizes[word];
	return 0;
}

// ---- host-only view of the 2-byte device layout of one row (tests): see usb_hostindex.h HostHalf
extern "C" int usb_debug_half_row(const uint32_t *targets, uint32_t n, uint32_t n_targets, uint16_t *out, uint32_t out_cap,
  uint32_t *groups, uint32_t *dummy0)
{
	if ((!targets && n) || !out || !groups || !dummy0)
		return fail(USB_EINVAL, "usb_debug_half_row: null argument");
	if (n_targets == 0 || n_targets > USB_HALF_MAX_TARGETS)
		return fail(USB_EINVAL, "usb_debug_half_row: %u targets (1..%u)", n_targets, USB_HALF_MAX_TARGETS);
	HostCSR H;
	H.base = 0;
	H.count = n_targets;
	H.slots = 1;
	H.row_off = {0, (uint64_t)((n + 3) & ~3u)};
	H.row_size = {n};
	H.postings.assign(((size_t)n + 3 & ~(size_t)3) + 4, 0xffffffffu);
	for (uint32_t i = 0; i < n; ++i) {
		if (targets[i] >= n_targets || (i && targets[i] <= targets[i - 1]))
			return fail(USB_EINVAL, "usb_debug_half_row: targets must be ascending and < n_targets");
		H.postings[i] = targets[i];
	}
	H.n_postings = n;
	HostHalf hh;
	make_half(H, n_targets, 1, hh);
	*groups = hh.row_groups[0];
	*dummy0 = hh.dummy0;
	if ((uint64_t)hh.row_groups[0] * 256 > out_cap)
		return fail(USB_ELIMIT, "usb_debug_half_row: %u entries needed", hh.row_groups[0] * 256);
	memcpy(out, hh.postings.data(), (size_t)hh.row_groups[0] * 256 * 2);
	return 0;
}
