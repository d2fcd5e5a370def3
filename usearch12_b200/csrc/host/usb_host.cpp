// usb_host.cpp -- see usb_host.h
#include "usb_host.h"
#include "../usb_tables.h"

#include <algorithm>
#include <cctype>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <memory>
#include <mutex>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <unordered_map>

namespace usbhost {

void Die(const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	fprintf(stderr, "\n---Fatal error---\n");
	vfprintf(stderr, fmt, ap);
	fprintf(stderr, "\n");
	va_end(ap);
	exit(1);
}

void Warning(const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	fprintf(stderr, "\nWARNING: ");
	vfprintf(stderr, fmt, ap);
	fprintf(stderr, "\n");
	va_end(ap);
}

static void CheckUsb(int rc, const char *what)
{
	if (rc != 0)
		Die("%s: %s (usb200 error %d)", what, usb_last_error(), rc);
}

// ------------------------------------------------------------------ SeqDB
// fastaseqsource.cpp:25-124: label = everything after '>', letters = isalpha characters, white
// space skipped, gap characters stripped, other bytes reported and skipped, empty sequences
// dropped with a warning.
namespace {
// One piece of a FASTA file (it starts at a '>' line, or at the top of the file) parsed on its own.
struct FastaPiece {
	std::vector<uint8_t> letters;
	std::vector<uint64_t> offsets{0};
	std::vector<std::string> labels;
	std::vector<std::pair<unsigned, std::string>> empty; // (line inside the piece, label) of empty sequences
	unsigned lines = 0, bad_bytes = 0;
	bool no_label = false;   // letters before the first '>' (only an error at the top of the file)
	unsigned no_label_line = 0;
};

void ParseFastaPiece(const char *p, const char *end, FastaPiece &P)
{
	P.letters.reserve((size_t)(end - p));
	bool have_label = false;
	std::string label;
	uint64_t start = 0;
	auto close_record = [&]() {
		if (!have_label)
			return;
		if (P.letters.size() > start) {
			P.labels.push_back(label);
			P.offsets.push_back(P.letters.size());
		} else
			P.empty.emplace_back(P.lines, label);
		start = P.letters.size();
	};
	while (p < end) {
		const char *eol = (const char *)memchr(p, '\n', (size_t)(end - p));
		if (!eol)
			eol = end;
		const char *q = eol;
		while (q > p && (q[-1] == '\r' || q[-1] == '\n'))
			--q;
		++P.lines;
		if (q > p && *p == '>') {
			close_record();
			label.assign(p + 1, q);
			have_label = true;
		} else if (q > p) {
			if (!have_label) {
				if (!P.no_label) {
					P.no_label = true;
					P.no_label_line = P.lines;
				}
			} else {
				// lines of letters only (the normal case) are appended in one piece
				const char *c = p;
				while (c < q && (((unsigned char)*c | 0x20u) - 'a') < 26u)
					++c;
				if (c == q)
					P.letters.insert(P.letters.end(), (const uint8_t *)p, (const uint8_t *)q);
				else
					for (c = p; c < q; ++c) {
						unsigned char ch = (unsigned char)*c;
						if (isalpha(ch))
							P.letters.push_back(ch);
						else if (isspace(ch) || ch == '-' || ch == '.')
							continue;
						else
							++P.bad_bytes;
					}
			}
		}
		p = eol + 1;
	}
	close_record();
}
} // namespace

// fastaseqsource.cpp:25-124 semantics (labels without '>', letters only, blank lines and white space
// skipped, empty sequences dropped with a warning).  Large files are cut at '>' lines and the pieces
// parsed by several threads; the result does not depend on the cut.
void SeqDB::FromFasta(const std::string &FileName)
{
	FILE *f = fopen(FileName.c_str(), "rb");
	if (!f)
		Die("Cannot open %s", FileName.c_str());
	fseek(f, 0, SEEK_END);
	long sz = ftell(f);
	fseek(f, 0, SEEK_SET);
	if (sz == 0)
		Die("Empty file %s", FileName.c_str()); // filetype.cpp:14-15
	// (not a std::vector: its zero fill touches every page once more, 0.15 s per 200 MB)
	std::unique_ptr<char[]> buf(new char[(size_t)sz + 1]);
	if (sz > 0 && fread(buf.get(), 1, (size_t)sz, f) != (size_t)sz)
		Die("Read error on %s", FileName.c_str());
	fclose(f);
	buf[sz] = '\n';
	const char *base = buf.get(), *end = buf.get() + sz;
	if (sz > 0 && base[0] == '@') { // filetype.cpp:20-29: the first byte decides between FASTA and FASTQ
		FromFastq(base, end, FileName);
		return;
	}
	unsigned T = 1;
	if (sz > (8 << 20))
		T = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
	if (const char *e = getenv("USB_FASTA_THREADS")) // test knob: any file, any number of pieces
		T = (unsigned)std::max(1, std::min(64, atoi(e)));
	// piece k starts at the first '>' at the beginning of a line at or after k * sz / T
	std::vector<const char *> cut(T + 1, end);
	cut[0] = base;
	for (unsigned k = 1; k < T; ++k) {
		const char *p = base + (size_t)sz * k / T;
		if (p < cut[k - 1])
			p = cut[k - 1];
		const char *hit = end;
		while (p < end) {
			const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
			if (!nl || nl + 1 >= end)
				break;
			if (nl[1] == '>') {
				hit = nl + 1;
				break;
			}
			p = nl + 1;
		}
		cut[k] = hit;
	}
	std::vector<FastaPiece> pieces(T);
	if (T == 1)
		ParseFastaPiece(base, end, pieces[0]);
	else {
		std::vector<std::thread> th;
		for (unsigned k = 0; k < T; ++k)
			th.emplace_back([&, k]() {
				if (cut[k] < cut[k + 1])
					ParseFastaPiece(cut[k], cut[k + 1], pieces[k]);
			});
		for (auto &t : th)
			t.join();
	}
	// messages in file order, with line numbers of the whole file
	unsigned line0 = 0, bad_bytes = 0;
	uint64_t total = 0;
	size_t n_seqs = 0;
	for (unsigned k = 0; k < T; ++k) {
		const FastaPiece &P = pieces[k];
		if (P.no_label) // only piece 0 can begin without a label: the others start at a '>' line
			Die("Bad FASTA file %s, expected '>' in line %u", FileName.c_str(), line0 + P.no_label_line);
		for (const auto &e : P.empty)
			Warning("Empty sequence at line %u in FASTA file %s, label >%s", line0 + e.first, FileName.c_str(), e.second.c_str());
		line0 += P.lines;
		bad_bytes += P.bad_bytes;
		total += P.letters.size();
		n_seqs += P.labels.size();
	}
	if (T == 1) {
		m_Letters.swap(pieces[0].letters);
		m_Offsets.swap(pieces[0].offsets);
		m_Labels.swap(pieces[0].labels);
	} else {
		m_Letters.resize(total);
		m_Offsets.assign(n_seqs + 1, 0);
		m_Labels.clear();
		m_Labels.resize(n_seqs);
		std::vector<uint64_t> l0(T + 1, 0);
		std::vector<size_t> s0(T + 1, 0);
		for (unsigned k = 0; k < T; ++k) {
			l0[k + 1] = l0[k] + pieces[k].letters.size();
			s0[k + 1] = s0[k] + pieces[k].labels.size();
		}
		std::vector<std::thread> th;
		for (unsigned k = 0; k < T; ++k)
			th.emplace_back([&, k]() {
				FastaPiece &P = pieces[k];
				if (!P.letters.empty())
					memcpy(m_Letters.data() + l0[k], P.letters.data(), P.letters.size());
				for (size_t i = 0; i < P.labels.size(); ++i) {
					m_Offsets[s0[k] + i + 1] = l0[k] + P.offsets[i + 1];
					m_Labels[s0[k] + i].swap(P.labels[i]);
				}
			});
		for (auto &t : th)
			t.join();
	}
	if (bad_bytes)
		Warning("%u invalid bytes in FASTA file %s ignored", bad_bytes, FileName.c_str());
}

// fastqseqsource.cpp:8-115: four lines per record -- "@label", letters, "+anything", qualities of the same
// length; carriage returns are dropped wherever they stand (linereader.cpp:116-117); empty lines are only
// allowed at the end of the file.  A record without letters is dropped with a warning, as in FASTA input.
namespace {
struct FastqPiece {
	std::vector<uint8_t> letters;
	std::vector<char> quals;
	std::vector<uint64_t> offsets{0};
	std::vector<std::string> labels;
	std::vector<std::pair<unsigned, std::string>> empty; // (line inside the piece, label) of records without letters
	unsigned lines = 0;
};

// Parses [p, end); a malformed record stops the program with the reference's message.
void ParseFastqPiece(const char *p, const char *end, FastqPiece &P, const char *fn)
{
	std::string crbuf, label;
	const char *lb = nullptr, *le = nullptr; // the current line without its line end
	unsigned &line_nr = P.lines;
	// LineReader::ReadLine (linereader.cpp:90-133): false at the end of the file; a last line without '\n' counts;
	// carriage returns are dropped wherever they stand (a line that has one is copied, the others are used in place)
	auto read_line = [&]() -> bool {
		if (p >= end)
			return false;
		const char *eol = (const char *)memchr(p, '\n', (size_t)(end - p));
		if (!eol)
			eol = end;
		lb = p;
		le = eol;
		p = eol + 1;
		if (le > lb && memchr(lb, '\r', (size_t)(le - lb))) {
			crbuf.assign(lb, le);
			crbuf.erase(std::remove(crbuf.begin(), crbuf.end(), '\r'), crbuf.end());
			lb = crbuf.data();
			le = lb + crbuf.size();
		}
		++line_nr;
		return true;
	};
#define FASTQ_FAIL(...) Die(__VA_ARGS__)
	P.letters.reserve((size_t)(end - p) / 2);
	P.quals.reserve((size_t)(end - p) / 2);
	while (read_line()) {
		if (lb == le) {
			for (;;) {
				const unsigned nr = line_nr;
				if (!read_line())
					return;
				if (lb != le)
					FASTQ_FAIL("Empty line nr %u in FASTQ file '%s'", nr, fn);
			}
		}
		if (*lb != '@')
			FASTQ_FAIL("Bad line %u in FASTQ file '%s': expected '@'", line_nr, fn);
		label.assign(lb + 1, le);
		if (!read_line())
			FASTQ_FAIL("Unexpected end-of-file in FASTQ file %s", fn);
		for (const char *c = lb; c < le; ++c)
			if ((((unsigned char)*c | 0x20u) - 'a') >= 26u) { // not a letter
				const unsigned char ch = (unsigned char)*c;
				if (isprint(ch))
					FASTQ_FAIL("Invalid sequence letter '%c' in FASTQ, line %u file %s", ch, line_nr, fn);
				FASTQ_FAIL("Non-printing byte 0x%02x in FASTQ sequence line %u file %s label %s", ch, line_nr, fn, label.c_str());
			}
		const size_t L = (size_t)(le - lb);
		P.letters.insert(P.letters.end(), (const uint8_t *)lb, (const uint8_t *)le);
		read_line(); // "+[label]": contents ignored
		if (!read_line())
			FASTQ_FAIL("Unexpected end-of-file in FASTQ file %s", fn);
		if ((size_t)(le - lb) != L)
			FASTQ_FAIL("Bad FASTQ record: %u bases, %u quals line %u file %s label %s", (unsigned)L, (unsigned)(le - lb), line_nr, fn,
			           label.c_str());
		P.quals.insert(P.quals.end(), lb, le);
		if (L == 0) {
			P.empty.emplace_back(line_nr - 2, label);
			continue;
		}
		P.labels.push_back(label);
		P.offsets.push_back(P.letters.size());
	}
#undef FASTQ_FAIL
}
} // namespace

// One pass over the file.  (Cutting the file at record starts and parsing the pieces with several threads, as the
// FASTA reader does, was measured and dropped: 0.60 s against 0.42 s for 400 000 reads of 250 letters on 8 cores --
// the merge copies letters and qualities once more and costs more than the parsing it spreads.)
void SeqDB::FromFastq(const char *base, const char *end, const std::string &FileName)
{
	const char *fn = FileName.c_str();
	FastqPiece P;
	ParseFastqPiece(base, end, P, fn);
	m_HasQual = true;
	for (const auto &e : P.empty)
		Warning("Empty sequence at line %u in FASTQ file %s, label @%s", e.first, fn, e.second.c_str());
	m_Letters.swap(P.letters);
	m_Quals.swap(P.quals);
	m_Offsets.swap(P.offsets);
	m_Labels.swap(P.labels);
}

void SeqDB::DropSmallerThan(unsigned MinSize)
{
	if (MinSize == 0)
		return;
	const uint32_t n = GetSeqCount();
	uint32_t kept = 0;
	uint64_t w = 0;
	for (uint32_t i = 0; i < n; ++i) {
		if (OtuTabSink::GetSizeFromLabel(m_Labels[i], 0xffffffffu) < MinSize)
			continue;
		const uint64_t b = m_Offsets[i], L = m_Offsets[i + 1] - b;
		if (w != b) {
			memmove(m_Letters.data() + w, m_Letters.data() + b, L);
			if (m_HasQual)
				memmove(m_Quals.data() + w, m_Quals.data() + b, L);
		}
		if (kept != i)
			m_Labels[kept].swap(m_Labels[i]);
		m_Offsets[kept] = w;
		w += L;
		++kept;
	}
	m_Offsets[kept] = w;
	m_Offsets.resize(kept + 1);
	m_Labels.resize(kept);
	m_Letters.resize(w);
	if (m_HasQual)
		m_Quals.resize(w);
}

void SeqDB::FromUDB(const std::string &FileName, bool &IsNucleo, uint32_t &WordLength)
{
	usb_udb *u = nullptr;
	CheckUsb(usb_udb_read(FileName.c_str(), &u), "usb_udb_read");
	const uint32_t n = usb_udb_seq_count(u);
	const uint64_t *off = nullptr;
	const uint8_t *letters = usb_udb_seqs(u, &off);
	m_Letters.assign(letters, letters + off[n]);
	m_Offsets.assign(off, off + n + 1);
	m_Labels.clear();
	for (uint32_t i = 0; i < n; ++i)
		m_Labels.push_back(usb_udb_label(u, i));
	IsNucleo = usb_udb_is_nucleo(u) != 0;
	WordLength = usb_udb_word_length(u);
	usb_udb_free(u);
}

bool IsUDBFile(const std::string &FileName) { return usb_udb_probe(FileName.c_str()) != 0; }

bool UDBIsNucleo(const std::string &FileName)
{
	usb_udb *u = nullptr;
	CheckUsb(usb_udb_read(FileName.c_str(), &u), "usb_udb_read");
	const bool nucleo = usb_udb_is_nucleo(u) != 0;
	usb_udb_free(u);
	return nucleo;
}

void MakeUDB(const std::string &FastaFileName, const std::string &OutputFileName, const usb_params &P)
{
	if (FastaFileName.empty() || OutputFileName.empty())
		Die("Missing input or output filename"); // makeudb.cpp:30-31
	SeqDB DB;
	DB.FromFasta(FastaFileName);
	std::vector<const char *> labels(DB.GetSeqCount());
	for (uint32_t i = 0; i < DB.GetSeqCount(); ++i)
		labels[i] = DB.GetLabel(i);
	CheckUsb(usb_udb_write(OutputFileName.c_str(), &P, DB.Letters(), DB.Offsets(), labels.data(), DB.GetSeqCount()), "usb_udb_write");
}

void SeqDB::GetSI(uint32_t Index, SeqInfo &SI) const
{
	SI.m_Label = m_Labels[Index].c_str();
	SI.m_Seq = GetSeq(Index);
	SI.m_L = GetSeqLength(Index);
	SI.m_Index = Index;
	SI.m_RevComp = false;
	SI.m_Qual = m_HasQual ? m_Quals.data() + m_Offsets[Index] : nullptr;
}

// ------------------------------------------------------------------ AlignResult
unsigned AlignResult::GetPathLength() const
{
	unsigned n = 0;
	for (uint32_t k = 0; k < m_Hit.run_cnt; ++k)
		n += m_Runs[k] >> 2;
	return n;
}

void AlignResult::GetPath(std::string &Path) const
{
	static const char ops[4] = {'M', 'D', 'I', '?'};
	Path.clear();
	for (uint32_t k = 0; k < m_Hit.run_cnt; ++k)
		Path.append(m_Runs[k] >> 2, ops[m_Runs[k] & 3]);
}

void AlignResult::GetCompressedPath(std::string &CPath) const
{
	static const char ops[4] = {'M', 'D', 'I', '?'};
	CPath.clear();
	char tmp[16];
	for (uint32_t k = 0; k < m_Hit.run_cnt; ++k) {
		const unsigned n = m_Runs[k] >> 2;
		if (n != 1) {
			snprintf(tmp, sizeof tmp, "%u", n);
			CPath += tmp;
		}
		CPath += ops[m_Runs[k] & 3];
	}
}

// ------------------------------------------------------------------ OutputSink
enum UserField {
	UF_query, UF_target, UF_clusternr, UF_id, UF_fractid, UF_dist, UF_pairs, UF_gaps, UF_allgaps, UF_qlo, UF_qhi,
	UF_tlo, UF_thi, UF_qlot, UF_qhit, UF_qunt, UF_tlot, UF_thit, UF_tunt, UF_ql, UF_tl, UF_alnlen, UF_opens,
	UF_exts, UF_aln, UF_caln, UF_qstrand, UF_tstrand, UF_mism, UF_ids, UF_diffs, UF_evalue, UF_bits, UF_raw, UF_qlor,
	UF_qhir, UF_tlor, UF_thir,
	// the fields that read the letters (rows, segments, substitution scores) and the derived ratios
	UF_mid, UF_pctpv, UF_pctgaps, UF_pv, UF_qs, UF_ts, UF_qseq, UF_tseq, UF_qseg, UF_tseg, UF_qsegf, UF_qrow, UF_trow,
	UF_qrowdots, UF_trowdots, UF_qframe, UF_tframe, UF_qcov, UF_tcov, UF_diffsa, UF_editdiffs, UF_abskew, UF_orflo,
	UF_orfhi, UF_orfframe, UF_gc, UF_kmerid, UF_qtrimlo, UF_qtrimhi, UF_qtrimseq, UF_COUNT
};
static const char *g_UserFieldNames[UF_COUNT] = {
	"query", "target", "clusternr", "id", "fractid", "dist", "pairs", "gaps", "allgaps", "qlo", "qhi", "tlo",
	"thi", "qlot", "qhit", "qunt", "tlot", "thit", "tunt", "ql", "tl", "alnlen", "opens", "exts", "aln", "caln",
	"qstrand", "tstrand", "mism", "ids", "diffs", "evalue", "bits", "raw", "qlor", "qhir", "tlor", "thir",
	"mid", "pctpv", "pctgaps", "pv", "qs", "ts", "qseq", "tseq", "qseg", "tseg", "qsegf", "qrow", "trow",
	"qrowdots", "trowdots", "qframe", "tframe", "qcov", "tcov", "diffsa", "editdiffs", "abskew", "orflo",
	"orfhi", "orfframe", "gc", "kmerid", "qtrimlo", "qtrimhi", "qtrimseq"};

// g_MatchMxNucleo / g_MatchMxAmino (alpha2.cpp:220-279) and g_SubstMx by raw character, from the
// same table builder the kernels use (usb_tables.h)
struct FormatTables {
	usb::LocalTables L;
	bool Match(uint8_t a, uint8_t b) const { return (L.match[L.code[a]] >> L.code[b]) & 1; }
	int Score(uint8_t a, uint8_t b) const { return L.score[L.code[a]][L.code[b]]; }
};

const uint8_t *AlignResult::GetQSeq(std::string &Buf) const
{
	if (!m_Hit.strand)
		return m_Query.m_Seq;
	static const usb::CharTables *T = []() {
		usb::CharTables *t = new usb::CharTables;
		usb::build_char_tables(*t);
		return t;
	}();
	Buf.resize(m_Query.m_L);
	for (uint32_t i = 0; i < m_Query.m_L; ++i)
		Buf[i] = (char)T->comp[m_Query.m_Seq[m_Query.m_L - 1 - i]];
	return (const uint8_t *)Buf.data();
}

namespace {
// The columns between the first and the last M of an alignment with the letters under them: what
// AlignResult::GetQueryRow / GetTargetRow / GetAnnotRow / Get*RowDots / GetPositiveCount / FillLo's
// m_DiffCountA / GetKmerId walk (arscorer.cpp:201-296,305-564,882-930).
struct RowWalk {
	std::string path, qbuf;
	const uint8_t *Q = nullptr, *T = nullptr; // first letters under the first M column
	size_t c0 = 0, c1 = 0;                    // [c0, c1) = columns m_FirstMCol .. m_LastMCol
	explicit RowWalk(const AlignResult &AR)
	{
		AR.GetPath(path);
		c0 = AR.m_Hit.first_mcol;
		c1 = c0 + AR.m_Hit.alnlen;
		Q = AR.GetQSeq(qbuf) + AR.m_Hit.first_mq;
		T = AR.m_Target.m_Seq + AR.m_Hit.first_mt;
	}
	// fn(op, q, t): q / t = raw letter, or 0 where the row has a gap
	template <class F> void Each(F fn) const
	{
		const uint8_t *q = Q, *t = T;
		for (size_t c = c0; c < c1; ++c) {
			const char op = path[c];
			const uint8_t qc = (op == 'M' || op == 'D') ? *q++ : 0;
			const uint8_t tc = (op == 'M' || op == 'I') ? *t++ : 0;
			fn(op, qc, tc);
		}
	}
};

char AnnotSym(const FormatTables &FT, bool Nucleo, uint8_t a, uint8_t b) // arscorer.cpp:12-46
{
	if (Nucleo) {
		if (toupper(a) == toupper(b) && strchr("ACGTUacgtu", a) && strchr("ACGTUacgtu", b))
			return '|';
		return FT.Match(a, b) ? '+' : ' ';
	}
	if (FT.Match(a, b))
		return '|';
	const int s = FT.Score(a, b);
	return s >= 2 ? ':' : s > 0 ? '.' : ' ';
}

void QueryRow(const RowWalk &W, std::string &Row)
{
	Row.clear();
	W.Each([&](char, uint8_t q, uint8_t) { Row += q ? (char)toupper(q) : '-'; });
}

void TargetRow(const RowWalk &W, std::string &Row)
{
	Row.clear();
	W.Each([&](char, uint8_t, uint8_t t) { Row += t ? (char)toupper(t) : '-'; });
}

void AnnotRow(const RowWalk &W, const FormatTables &FT, bool Nucleo, std::string &Row)
{
	Row.clear();
	W.Each([&](char op, uint8_t q, uint8_t t) { Row += op == 'M' ? AnnotSym(FT, Nucleo, q, t) : ' '; });
}

// rows with '.' where the upper-cased letters match (a letter never matches the gap symbol)
void RowDots(const RowWalk &W, const FormatTables &FT, bool QueryRowWanted, std::string &Row)
{
	Row.clear();
	W.Each([&](char, uint8_t q, uint8_t t) {
		const uint8_t uq = q ? (uint8_t)toupper(q) : (uint8_t)'-', ut = t ? (uint8_t)toupper(t) : (uint8_t)'-';
		const uint8_t mine = QueryRowWanted ? uq : ut;
		const bool have = QueryRowWanted ? q != 0 : t != 0;
		Row += !have ? '-' : FT.Match(uq, ut) ? '.' : (char)mine;
	});
}

unsigned PositiveCount(const RowWalk &W, const FormatTables &FT)
{
	unsigned n = 0;
	W.Each([&](char op, uint8_t q, uint8_t t) { n += op == 'M' && FT.Score(q, t) > 0; });
	return n;
}

unsigned DiffCountA(const RowWalk &W)
{
	unsigned n = 0;
	W.Each([&](char op, uint8_t q, uint8_t t) { n += op == 'M' && toupper(q) != toupper(t); });
	return n;
}

double KmerId(const AlignResult &AR, const RowWalk &W, unsigned w)
{
	const unsigned MinL = std::min(AR.m_Hit.ql, AR.m_Hit.tl);
	if (MinL < w)
		return 0.0;
	unsigned run = 0, matches = 0;
	W.Each([&](char op, uint8_t q, uint8_t t) {
		run = (op == 'M' && toupper(q) == toupper(t)) ? run + 1 : 0;
		matches += op == 'M' && run >= w;
	});
	return double(matches) / double(MinL - w + 1);
}

// SeqToFasta (seqdb.cpp:62-95) with the default -fasta_cols 80
void AppendFasta80(std::string &out, const char *Label, const uint8_t *Seq, size_t L)
{
	if (L == 0)
		return;
	out += '>';
	out += Label;
	out += '\n';
	for (size_t i = 0; i < L; i += 80) {
		out.append((const char *)Seq + i, std::min<size_t>(80, L - i));
		out += '\n';
	}
}

// RowToFasta (outputsink.cpp:30-58): the letters of a row; an empty row still gets its newline
void AppendRowFasta(std::string &out, const char *Label, const std::string &Row)
{
	out += '>';
	out += Label;
	unsigned n = 0;
	for (char c : Row) {
		if (c == '-' || c == '.')
			continue;
		if (n % 80 == 0)
			out += '\n';
		out += c;
		++n;
	}
	out += '\n';
}

// AlignResult::GetTrimInfo (arscorer.cpp:932-970): the query without the letters under a leading / trailing run of D
// columns; the loop that copies the letters stops one short of QHi, as in the reference
void TrimInfo(const AlignResult &AR, unsigned &QLo, unsigned &QHi, std::string &QSeg)
{
	const unsigned QL = AR.m_Hit.ql;
	QLo = 0;
	QHi = QL ? QL - 1 : 0;
	QSeg.clear();
	if (QL == 0)
		return;
	const uint32_t n = AR.m_Hit.run_cnt;
	if (n > 0 && (AR.m_Runs[0] & 3) == 1)
		QLo = AR.m_Runs[0] >> 2;
	if (n > 0 && (AR.m_Runs[n - 1] & 3) == 1) {
		const unsigned NewQHi = QL - (AR.m_Runs[n - 1] >> 2) - 1;
		if (NewQHi > QLo)
			QHi = NewQHi;
	}
	std::string buf;
	const uint8_t *Q = AR.GetQSeq(buf);
	if (QHi > QLo)
		QSeg.assign((const char *)Q + QLo, QHi - QLo);
}

unsigned NDig(unsigned n) // alnout.cpp:9-24
{
	return n < 10 ? 1 : n < 100 ? 2 : n < 1000 ? 3 : n < 10000 ? 4 : n < 100000 ? 5 : n < 1000000 ? 6 : 10;
}
} // namespace

OutputSink::OutputSink(const OutputOpts &O) : m_O(O)
{
	auto open = [](const std::string &fn) -> FILE * {
		if (fn.empty())
			return nullptr;
		FILE *f = fopen(fn.c_str(), "wb");
		if (!f)
			Die("Cannot create %s", fn.c_str());
		return f;
	};
	m_f[O_UC] = open(O.uc);
	m_f[O_B6] = open(O.blast6out);
	m_f[O_USER] = open(O.userout);
	m_f[O_ALN] = open(O.alnout);
	m_f[O_PAIRS] = open(O.fastapairs);
	m_f[O_QSEG] = open(O.qsegout);
	m_f[O_TSEG] = open(O.tsegout);
	m_f[O_MATCHED] = open(O.matched);
	m_f[O_NOTMATCHED] = open(O.notmatched);
	m_f[O_TRIM] = open(O.trimout);
	m_f[O_MATCHEDFQ] = open(O.matchedfq);
	m_f[O_NOTMATCHEDFQ] = open(O.notmatchedfq);
	m_OutputNoHits = O.output_no_hits;
	m_T = std::make_shared<FormatTables>();
	usb::build_local_tables(O.nucleo, O.match, O.mismatch, m_T->L);
	if (m_f[O_ALN]) {
		// outputsink.cpp:142-147: the command line, then the program line (version, memory and cores there)
		fprintf(m_f[O_ALN], "%s\nusearch12_b200 (B200 hot path of usearch v12.0)\n", O.cmdline.c_str());
	}
	if (m_f[O_USER]) {
		// outputsink.cpp:149-156: -userout needs -userfields; userout.cpp:20-60: fields separated by '+'
		if (O.userfields.empty())
			Die("--userout requires --userfields");
		const std::string &spec = O.userfields;
		size_t pos = 0;
		while (pos <= spec.size()) {
			size_t e = spec.find('+', pos);
			if (e == std::string::npos)
				e = spec.size();
			std::string name = spec.substr(pos, e - pos);
			int idx = -1;
			for (int i = 0; i < UF_COUNT; ++i)
				if (name == g_UserFieldNames[i])
					idx = i;
			if (idx < 0)
				Die("Invalid or unsupported userfield name '%s'", name.c_str());
			// For a global alignment the reference prints m_HSP.Leni letters from the first M position
			// (alignresult.h:173, userout.cpp:204-207, arscorer.cpp:863-880): past the end of the sequence as soon
			// as the alignment starts with a terminal gap.  Nothing to be identical to.
			if (!O.local && (idx == UF_qseg || idx == UF_tseg || idx == UF_gc))
				Die("userfield %s is only supported with -usearch_local (the reference reads past the end of the "
				    "sequence for global alignments)", name.c_str());
			m_UserFields.push_back(idx);
			pos = e + 1;
		}
	}
}

OutputSink::~OutputSink() { OnAllDone(); }

void OutputSink::Flush(int k, std::string &buf, bool force)
{
	if (m_f[k] && (force || buf.size() > (1u << 20))) {
		fwrite(buf.data(), 1, buf.size(), m_f[k]);
		buf.clear();
	}
}

void OutputSink::OnAllDone()
{
	for (int k = 0; k < O_COUNT; ++k) {
		Flush(k, m_b[k], true);
		if (m_f[k]) {
			fclose(m_f[k]);
			m_f[k] = nullptr;
		}
	}
}

static void appendf(std::string &s, const char *fmt, ...)
{
	char tmp[256];
	va_list ap;
	va_start(ap, fmt);
	int n = vsnprintf(tmp, sizeof tmp, fmt, ap);
	va_end(ap);
	if (n > 0)
		s.append(tmp, (size_t)std::min<int>(n, (int)sizeof tmp - 1));
}

// outputuc.cpp:19-22,45-69
void OutputSink::OutputUC(const SeqInfo &Query, const HitMgr &HM, std::string &m_bUC) const
{
	if (!m_f[O_UC])
		return;
	std::string cp;
	if (HM.m_Hits.empty() && !m_O.uc_hitsonly) {
		appendf(m_bUC, "N\t*\t%u\t*\t.\t*\t*\t*\t", Query.m_L);
		m_bUC += Query.m_Label;
		m_bUC += "\t*\n";
	}
	for (const AlignResult &AR : HM.m_Hits) {
		AR.GetCompressedPath(cp);
		appendf(m_bUC, "H\t%u\t%u\t%.1f\t%c\t%u\t%u\t", AR.GetTargetIndex(), AR.GetIQL(), AR.GetPctId(),
		  AR.GetQueryStrand(), AR.GetIQLo(), AR.GetITLo());
		m_bUC += cp;
		m_bUC += '\t';
		m_bUC += AR.GetQueryLabel();
		m_bUC += '\t';
		m_bUC += AR.GetTargetLabel();
		m_bUC += '\n';
	}
}

// blast6out.cpp:27-80 (global alignments: evalue and bit score print as '*')
void OutputSink::OutputBlast6(const HitMgr &HM, std::string &m_bB6) const
{
	if (!m_f[O_B6])
		return;
	if (HM.m_Hits.empty() && m_OutputNoHits) { // blast6out.cpp:82-103
		m_bB6 += HM.m_Query.m_Label;
		m_bB6 += "\t*\t0\t0\t0\t0\t0\t0\t0\t0\t*\t0\n";
	}
	for (const AlignResult &AR : HM.m_Hits) {
		m_bB6 += AR.GetQueryLabel();
		m_bB6 += '\t';
		m_bB6 += AR.GetTargetLabel();
		appendf(m_bB6, "\t%.1f\t%u\t%u\t%u\t%u\t%u\t%u\t%u", AR.GetPctId(), AR.GetAlnLength(),
		  AR.GetMismatchCount(), AR.GetGapOpenCount(), AR.GetIQLo1(), AR.GetIQHi1(), AR.GetTLo6(), AR.GetTHi6());
		if (AR.IsLocal())
			appendf(m_bB6, "\t%.2g\t%.1f\n", AR.GetEvalue(), AR.GetBitScore());
		else
			m_bB6 += "\t*\t*\n";
	}
}

// userout.cpp:126-352
void OutputSink::OutputUser(const HitMgr &HM, std::string &m_bUser) const
{
	if (!m_f[O_USER])
		return;
	std::string tmp;
	if (HM.m_Hits.empty() && m_OutputNoHits) { // userout.cpp:53-124
		for (size_t i = 0; i < m_UserFields.size(); ++i) {
			if (i)
				m_bUser += '\t';
			switch (m_UserFields[i]) {
			case UF_query: m_bUser += HM.m_Query.m_Label; break;
			case UF_ql: appendf(m_bUser, "%u", HM.m_Query.m_L); break;
			case UF_clusternr: appendf(m_bUser, "%u", 0xffffffffu); break; // hitmgr.cpp:84 m_QueryClusterIndex = UINT_MAX
			case UF_target: case UF_evalue: case UF_id: case UF_fractid: case UF_pairs: case UF_gaps: case UF_qlo: case UF_qhi:
			case UF_qlor: case UF_qhir: case UF_tlo: case UF_thi: case UF_tlor: case UF_thir: case UF_tl: case UF_alnlen:
			case UF_opens: case UF_exts: case UF_raw: case UF_bits: case UF_aln: case UF_caln: case UF_qstrand:
			case UF_tstrand: case UF_mism: case UF_ids: case UF_diffs:
			case UF_mid: case UF_pctpv: case UF_pctgaps: case UF_pv: case UF_qs: case UF_ts: case UF_qrow: case UF_trow:
			case UF_qframe: case UF_tframe: case UF_qcov: case UF_tcov: case UF_diffsa: case UF_abskew: case UF_tseq:
				m_bUser += '*';
				break;
			case UF_qseq: m_bUser.append((const char *)HM.m_Query.m_Seq, HM.m_Query.m_L); break;
			default:
				Die("Invalid user field index %u (-output_no_hits)", (unsigned)m_UserFields[i]);
			}
		}
		m_bUser += '\n';
	}
	for (const AlignResult &AR : HM.m_Hits) {
		std::unique_ptr<RowWalk> walk; // built on demand, once per hit
		auto W = [&]() -> const RowWalk & {
			if (!walk)
				walk.reset(new RowWalk(AR));
			return *walk;
		};
		for (size_t i = 0; i < m_UserFields.size(); ++i) {
			if (i)
				m_bUser += '\t';
			switch (m_UserFields[i]) {
			case UF_query: m_bUser += AR.GetQueryLabel(); break;
			case UF_target: m_bUser += AR.GetTargetLabel(); break;
			case UF_clusternr: appendf(m_bUser, "%u", AR.GetTargetIndex()); break;
			case UF_id: appendf(m_bUser, "%.1f", AR.GetPctId()); break;
			case UF_fractid: appendf(m_bUser, "%.4f", AR.GetFractId()); break;
			case UF_dist: appendf(m_bUser, "%.4f", 1.0 - AR.GetFractId()); break;
			case UF_pairs: appendf(m_bUser, "%u", AR.GetLetterPairCount()); break;
			case UF_gaps: appendf(m_bUser, "%u", AR.GetGapCount()); break;
			case UF_allgaps: appendf(m_bUser, "%u", AR.GetAllGapCount()); break;
			case UF_qlo: appendf(m_bUser, "%u", AR.GetIQLo1()); break;
			case UF_qhi: appendf(m_bUser, "%u", AR.GetIQHi1()); break;
			case UF_tlo: appendf(m_bUser, "%u", AR.GetITLo1()); break;
			case UF_thi: appendf(m_bUser, "%u", AR.GetITHi1()); break;
			case UF_qlot: appendf(m_bUser, "%u", AR.GetQLoT()); break;
			case UF_qhit: appendf(m_bUser, "%u", AR.GetQHiT()); break;
			case UF_qunt: appendf(m_bUser, "%u", AR.GetQUnT()); break;
			case UF_tlot: appendf(m_bUser, "%u", AR.GetTLoT()); break;
			case UF_thit: appendf(m_bUser, "%u", AR.GetTHiT()); break;
			case UF_tunt: appendf(m_bUser, "%u", AR.GetTUnT()); break;
			case UF_ql: appendf(m_bUser, "%u", AR.GetIQL()); break;
			case UF_tl: appendf(m_bUser, "%u", AR.GetITL()); break;
			case UF_alnlen: appendf(m_bUser, "%u", AR.GetAlnLength()); break;
			case UF_opens: appendf(m_bUser, "%u", AR.GetGapOpenCount()); break;
			case UF_exts: appendf(m_bUser, "%u", AR.GetGapExtCount()); break;
			case UF_aln: AR.GetPath(tmp); m_bUser += tmp; break;
			case UF_caln: AR.GetCompressedPath(tmp); m_bUser += tmp; break;
			case UF_qstrand: m_bUser += AR.GetQueryStrand(); break;
			case UF_tstrand: m_bUser += AR.GetTargetStrand(); break;
			case UF_mism: appendf(m_bUser, "%u", AR.GetMismatchCount()); break;
			case UF_ids: appendf(m_bUser, "%u", AR.GetIdCount()); break;
			case UF_diffs: appendf(m_bUser, "%u", AR.GetDiffCount()); break;
			case UF_evalue: appendf(m_bUser, "%.3g", AR.GetEvalue()); break;
			case UF_bits: appendf(m_bUser, "%.0f", AR.GetBitScore()); break;
			case UF_raw: appendf(m_bUser, "%.0f", AR.GetRawScore()); break;
			case UF_qlor: appendf(m_bUser, "%u", AR.GetLoi()); break;
			case UF_qhir: appendf(m_bUser, "%u", AR.GetHii()); break;
			case UF_tlor: appendf(m_bUser, "%u", AR.GetLoj()); break;
			case UF_thir: appendf(m_bUser, "%u", AR.GetHij()); break;
			case UF_mid: appendf(m_bUser, "%.1f", 100.0 * AR.GetFractMatchId()); break;
			case UF_pctpv: {
				const unsigned L = AR.GetAlnLength();
				appendf(m_bUser, "%.1f", L == 0 ? 0.0 : 100.0 * (double(PositiveCount(W(), *m_T)) / double(L)));
				break;
			}
			case UF_pctgaps: appendf(m_bUser, "%.1f", AR.GetPctGaps()); break;
			case UF_pv: appendf(m_bUser, "%u", PositiveCount(W(), *m_T)); break;
			case UF_qs: appendf(m_bUser, "%u", AR.GetQuerySegLength()); break;
			case UF_ts: appendf(m_bUser, "%u", AR.GetTargetSegLength()); break;
			case UF_qseq: m_bUser.append((const char *)AR.GetQSeq(tmp), AR.GetIQL()); break;
			case UF_tseq: m_bUser.append((const char *)AR.m_Target.m_Seq, AR.GetITL()); break;
			case UF_qseg: m_bUser.append((const char *)W().Q, AR.GetQuerySegLength()); break;
			case UF_tseg: m_bUser.append((const char *)W().T, AR.GetTargetSegLength()); break;
			case UF_qsegf: { // userout.cpp:216-235: the query segment between '-' with up to -flank letters around it
				const unsigned f = m_O.flank, QL = AR.GetIQL(), Lo = AR.GetLoi(), Hi = AR.GetHii();
				const unsigned fl = std::min(Lo, f), fr = std::min(QL - Hi - 1, f);
				const char *Q = (const char *)AR.GetQSeq(tmp);
				m_bUser.append(Q + Lo - fl, fl);
				m_bUser += '-';
				m_bUser.append(Q + Lo, Hi - Lo + 1);
				m_bUser += '-';
				m_bUser.append(Q + Hi + 1, fr);
				break;
			}
			case UF_qrow: QueryRow(W(), tmp); m_bUser += tmp; break;
			case UF_trow: TargetRow(W(), tmp); m_bUser += tmp; break;
			case UF_qrowdots: RowDots(W(), *m_T, true, tmp); m_bUser += tmp; break;
			case UF_trowdots: RowDots(W(), *m_T, false, tmp); m_bUser += tmp; break;
			case UF_qframe: case UF_tframe: case UF_orfframe: m_bUser += "+0"; break; // no ORF queries on this path
			case UF_orflo: case UF_orfhi: m_bUser += '0'; break;
			case UF_qcov: appendf(m_bUser, "%.0f", 100.0 * AR.GetQueryCov()); break;
			case UF_tcov: appendf(m_bUser, "%.0f", 100.0 * AR.GetTargetCov()); break;
			case UF_diffsa: appendf(m_bUser, "%u", DiffCountA(W()) + 0u); break;
			case UF_editdiffs: appendf(m_bUser, "%u", AR.GetEditDiffCount()); break;
			case UF_abskew: { // arscorer.cpp:809-816
				const unsigned QSize = OtuTabSink::GetSizeFromLabel(AR.GetQueryLabel(), 0xffffffffu);
				const unsigned TSize = OtuTabSink::GetSizeFromLabel(AR.GetTargetLabel(), 0xffffffffu);
				appendf(m_bUser, "%.1f", double(TSize) / double(QSize));
				break;
			}
			case UF_gc: { // arscorer.cpp:863-880: C and G (either case) among the letters of the query segment
				const unsigned L = AR.GetQuerySegLength();
				unsigned n = 0;
				for (unsigned k = 0; k < L; ++k)
					n += strchr("CGcg", W().Q[k]) != nullptr;
				appendf(m_bUser, "%.1f", L == 0 ? 0.0 : (100.0 * n) / L);
				break;
			}
			case UF_kmerid: appendf(m_bUser, "%.4f", KmerId(AR, W(), m_O.wordlength)); break;
			case UF_qtrimlo: case UF_qtrimhi: case UF_qtrimseq: {
				unsigned Lo, Hi;
				TrimInfo(AR, Lo, Hi, tmp);
				if (m_UserFields[i] == UF_qtrimseq)
					m_bUser += tmp;
				else
					appendf(m_bUser, "%u", (m_UserFields[i] == UF_qtrimlo ? Lo : Hi) + 1);
				break;
			}
			}
		}
		m_bUser += '\n';
	}
}

// The per-query table at the top of a query's -alnout section (outputsink.cpp:237-356)
void OutputSink::OutputReport(const SeqInfo &Query, const HitMgr &HM, std::string &out) const
{
	if (!m_f[O_ALN] || HM.m_Hits.empty())
		return;
	out += "\nQuery >";
	out += Query.m_Label;
	out += '\n';
	if (!m_O.local) {
		out += " %Id   TLen  Target\n";
		for (const AlignResult &AR : HM.m_Hits) {
			appendf(out, "%3.0f%%  %5u  ", AR.GetPctId(), AR.GetITL());
			out += AR.GetTargetLabel();
			out += '\n';
		}
		return;
	}
	out += " Score     Evalue   %Id    QueryLo-Hi(Un)   TargetLo-Hi(Un)";
	if (m_O.nucleo)
		out += "  +";
	out += "  Target\n";
	for (const AlignResult &AR : HM.m_Hits) {
		char seg[64];
		appendf(out, "%6.0f  %9.1g  %3.0f%%", AR.GetRawScore(), AR.GetEvalue(), AR.GetPctId());
		snprintf(seg, sizeof seg, "%u-%u(%u)", AR.GetIQLo() + 1, AR.GetIQHi() + 1, AR.GetIQL() - AR.GetIQHi() - 1);
		appendf(out, "  %16s", seg);
		snprintf(seg, sizeof seg, "%u-%u(%u)", AR.GetITLo() + 1, AR.GetITHi() + 1, AR.GetITL() - AR.GetITHi() - 1);
		appendf(out, "  %16s", seg);
		if (m_O.nucleo)
			appendf(out, "  %c", AR.GetQueryStrand());
		out += "  ";
		out += AR.GetTargetLabel();
		out += '\n';
	}
}

// alnout.cpp:45-171 WriteAln
void OutputSink::OutputAln(const AlignResult &AR, std::string &out) const
{
	const RowWalk W(AR);
	std::string QRow, TRow, ARow;
	QueryRow(W, QRow);
	TargetRow(W, TRow);
	AnnotRow(W, *m_T, m_O.nucleo, ARow);
	const unsigned IQL = AR.GetIQL(), ITL = AR.GetITL();
	const unsigned w = NDig(std::max(IQL, ITL));
	const char *ntaa = m_O.nucleo ? "nt" : "aa";
	out += '\n';
	appendf(out, " Query %*u%s >", (int)w, IQL, ntaa);
	out += AR.GetQueryLabel();
	appendf(out, "\nTarget %*u%s >", (int)w, ITL, ntaa);
	out += AR.GetTargetLabel();
	out += "\n\n";
	const char QueryStrand = AR.GetQueryStrand(), TargetStrand = AR.GetTargetStrand();
	const bool ShowStrand = QueryStrand != '.';
	const unsigned AlnLength = (unsigned)QRow.size();
	const unsigned RowLen = m_O.rowlen;
	// positions of the letters at the two ends of each block; a block that is all gaps in one
	// sequence repeats the position without the +1 (alnout.cpp:26-43,104-126, alignresult.cpp:365-379)
	unsigned QPos = AR.m_Hit.first_mq, TPos = AR.m_Hit.first_mt;
	bool QAllGaps = false, TAllGaps = false;
	auto ipos_q = [&](unsigned Pos, bool AllGaps) { return (AR.m_Hit.strand ? IQL - Pos - 1 : Pos) + (AllGaps ? 0u : 1u); };
	auto ipos_t = [&](unsigned Pos, bool AllGaps) { return Pos + (AllGaps ? 0u : 1u); };
	auto advance = [](unsigned Pos, const char *Row, unsigned n, bool &AllGaps) {
		bool got = false;
		for (unsigned i = 0; i < n; ++i)
			if (Row[i] != '-') {
				if (got)
					++Pos;
				got = true;
			}
		AllGaps = !got;
		return Pos;
	};
	for (unsigned From = 0; From < AlnLength; From += RowLen) {
		const unsigned n = std::min(RowLen, AlnLength - From);
		const unsigned QFrom = ipos_q(QPos, QAllGaps), TFrom = ipos_t(TPos, TAllGaps);
		QPos = advance(QPos, QRow.data() + From, n, QAllGaps);
		TPos = advance(TPos, TRow.data() + From, n, TAllGaps);
		const unsigned QTo = ipos_q(QPos, QAllGaps), TTo = ipos_t(TPos, TAllGaps);
		if (!QAllGaps)
			++QPos;
		if (!TAllGaps)
			++TPos;
		appendf(out, "Qry %*u", (int)w, QFrom);
		if (ShowStrand)
			appendf(out, " %c", QueryStrand);
		out += ' ';
		out.append(QRow, From, n);
		appendf(out, " %u\n", QTo);
		out.append(4 + w + (ShowStrand ? 2 : 0) + 1, ' ');
		out.append(ARow, From, n);
		out += '\n';
		appendf(out, "Tgt %*u", (int)w, TFrom);
		if (ShowStrand)
			appendf(out, " %c", TargetStrand);
		out += ' ';
		out.append(TRow, From, n);
		appendf(out, " %u\n\n", TTo);
	}
	const unsigned Ids = AR.GetIdCount(), Gaps = AR.GetGapCount();
	appendf(out, "%u cols, %u ids (%.1f%%), %u gaps (%.1f%%)", AlnLength, Ids, AlnLength ? 100.0 * (double(Ids) / double(AlnLength)) : 0.0,
	  Gaps, AlnLength ? 100.0 * (double(Gaps) / double(AlnLength)) : 0.0);
	if (AR.IsLocal())
		appendf(out, ", score %.1f (%.1f bits), Evalue %.2g", AR.GetRawScore(), AR.GetBitScore(), AR.GetEvalue());
	out += '\n';
}

// OutputSink::OnQueryDone (outputsink.cpp:358-400): the report, every hit through every per-hit file,
// then the query into -matched or -notmatched
void OutputSink::FormatQuery(const SeqInfo &Query, const HitMgr &HM, Bufs &out) const
{
	OutputUC(Query, HM, out[O_UC]);
	OutputBlast6(HM, out[O_B6]);
	OutputUser(HM, out[O_USER]);
	OutputReport(Query, HM, out[O_ALN]);
	if (m_f[O_ALN] || m_f[O_PAIRS] || m_f[O_QSEG] || m_f[O_TSEG] || m_f[O_TRIM]) {
		std::string QRow, TRow;
		for (const AlignResult &AR : HM.m_Hits) {
			if (m_f[O_ALN])
				OutputAln(AR, out[O_ALN]);
			if (m_f[O_TRIM]) { // OutputSink::OutputTrim (outputsink.cpp:402-417): label:lo-hi
				unsigned Lo, Hi;
				std::string Seq;
				TrimInfo(AR, Lo, Hi, Seq);
				std::string Label = AR.GetQueryLabel();
				appendf(Label, ":%u-%u", Lo + 1, Hi + 1);
				AppendFasta80(out[O_TRIM], Label.c_str(), (const uint8_t *)Seq.data(), Seq.size());
			}
			if (!(m_f[O_PAIRS] || m_f[O_QSEG] || m_f[O_TSEG]))
				continue;
			const RowWalk W(AR);
			QueryRow(W, QRow);
			TargetRow(W, TRow);
			if (m_f[O_PAIRS]) { // outputsink.cpp:223-235
				std::string &o = out[O_PAIRS];
				o += '>';
				o += AR.GetQueryLabel();
				o += '\n';
				o += QRow;
				o += "\n>";
				o += AR.GetTargetLabel();
				o += '\n';
				o += TRow;
				o += "\n\n";
			}
			if (m_f[O_QSEG])
				AppendRowFasta(out[O_QSEG], AR.GetQueryLabel(), QRow);
			if (m_f[O_TSEG])
				AppendRowFasta(out[O_TSEG], AR.GetTargetLabel(), TRow);
		}
	}
	if (m_f[O_MATCHED] && !HM.m_Hits.empty())
		AppendFasta80(out[O_MATCHED], Query.m_Label, Query.m_Seq, Query.m_L);
	if (m_f[O_NOTMATCHED] && HM.m_Hits.empty())
		AppendFasta80(out[O_NOTMATCHED], Query.m_Label, Query.m_Seq, Query.m_L);
	const int fq = HM.m_Hits.empty() ? O_NOTMATCHEDFQ : O_MATCHEDFQ;
	if (m_f[fq]) { // SeqToFastq (seqdb.cpp:14-28)
		if (!Query.m_Qual)
			Die("Cannot convert FASTA to FASTQ");
		std::string &o = out[fq];
		o += '@';
		o += Query.m_Label;
		o += '\n';
		o.append((const char *)Query.m_Seq, Query.m_L);
		o += "\n+\n";
		o.append(Query.m_Qual, Query.m_L);
		o += '\n';
	}
}

void OutputSink::OnQueryDone(const SeqInfo &Query, const HitMgr &HM)
{
	FormatQuery(Query, HM, m_b);
	for (int k = 0; k < O_COUNT; ++k)
		Flush(k, m_b[k], false);
}

void OutputSink::OnBatchDone(const std::vector<HitMgr> &Batch)
{
	const size_t n = Batch.size();
	// queries per formatting thread; USB_FORMAT_CHUNK lowers it so that small tests reach the threaded path
	const char *env = getenv("USB_FORMAT_CHUNK");
	const size_t per = env && atoi(env) > 0 ? (size_t)atoi(env) : 4096;
	const unsigned T = (unsigned)std::max<size_t>(1, std::min<size_t>({(size_t)std::thread::hardware_concurrency(), (size_t)16, n / per}));
	if (T <= 1) {
		for (const HitMgr &HM : Batch)
			OnQueryDone(HM.m_Query, HM);
		return;
	}
	struct Chunk {
		Bufs b;
	};
	std::vector<Chunk> chunks(T);
	std::vector<std::thread> th;
	for (unsigned k = 0; k < T; ++k)
		th.emplace_back([&, k]() {
			Chunk &c = chunks[k];
			for (size_t i = n * k / T; i < n * (k + 1) / T; ++i)
				FormatQuery(Batch[i].m_Query, Batch[i], c.b);
		});
	for (auto &t : th)
		t.join();
	for (int k = 0; k < O_COUNT; ++k) {
		Flush(k, m_b[k], true);
		for (Chunk &c : chunks)
			Flush(k, c.b[k], true);
	}
}

// ------------------------------------------------------------------ DBHitSink
DBHitSink::DBHitSink(const SeqDB &DB, const std::string &DbMatched, const std::string &DbNotMatched, bool SizeIn, bool SizeOut,
  const std::string &DbCutOut, bool TopHitOnly)
  : m_DB(DB), m_DbMatched(DbMatched), m_DbNotMatched(DbNotMatched), m_DbCutOut(DbCutOut), m_SizeIn(SizeIn), m_SizeOut(SizeOut),
    m_TopHitOnly(TopHitOnly), m_HitCounts(DB.GetSeqCount(), 0)
{
	if (!m_DbCutOut.empty()) {
		m_Los.resize(DB.GetSeqCount());
		m_His.resize(DB.GetSeqCount());
	}
}

DBHitSink::~DBHitSink() { OnAllDone(); }

void DBHitSink::OnQueryDone(const SeqInfo &Query, const HitMgr &HM) // dbhitsink.cpp:132-159
{
	if (HM.m_Hits.empty())
		return;
	const unsigned N = m_SizeIn ? OtuTabSink::GetSizeFromLabel(Query.m_Label, 1) : 1;
	for (const AlignResult &AR : HM.m_Hits) {
		m_HitCounts[AR.GetTargetIndex()] += N;
		if (!m_DbCutOut.empty()) { // GetTLo / GetTHi: first and last target letter under an M column
			m_Los[AR.GetTargetIndex()].insert(m_Los[AR.GetTargetIndex()].end(), N, AR.m_Hit.first_mt);
			m_His[AR.GetTargetIndex()].insert(m_His[AR.GetTargetIndex()].end(), N, AR.m_Hit.last_mt);
		}
		if (m_TopHitOnly)
			break;
	}
}

void DBHitSink::OnAllDone() // dbhitsink.cpp:42-50,108-130
{
	if (m_Done)
		return;
	m_Done = true;
	std::string Stored;
	if (!m_DbCutOut.empty()) {
		// CutToFASTA (dbhitsink.cpp:52-106): every target with hits, from the median first to the median last aligned
		// position of its hits (element N/2 of the sorted positions)
		FILE *f = fopen(m_DbCutOut.c_str(), "wb");
		if (!f)
			Die("Cannot create %s", m_DbCutOut.c_str());
		std::string out;
		for (uint32_t i = 0; i < m_DB.GetSeqCount(); ++i) {
			if (m_HitCounts[i] == 0)
				continue;
			std::sort(m_Los[i].begin(), m_Los[i].end());
			std::sort(m_His[i].begin(), m_His[i].end());
			const unsigned Lo = m_Los[i][m_Los[i].size() / 2], Hi = m_His[i][m_His[i].size() / 2];
			if (!(Lo < Hi && Hi < m_DB.GetSeqLength(i)))
				Die("-dbcutout: segment %u-%u of >%s", Lo, Hi, m_DB.GetLabel(i)); // asserta(Lo < Hi && Hi < L)
			const uint8_t *Seq = m_DB.GetSeq(i);
			if (m_Searcher) {
				m_Searcher->GetStoredTarget(i, Stored);
				Seq = (const uint8_t *)Stored.data();
			}
			AppendFasta80(out, m_DB.GetLabel(i), Seq + Lo, Hi - Lo + 1);
		}
		fwrite(out.data(), 1, out.size(), f);
		fclose(f);
	}
	for (int Matched = 1; Matched >= 0; --Matched) {
		const std::string &FileName = Matched ? m_DbMatched : m_DbNotMatched;
		if (FileName.empty())
			continue;
		FILE *f = fopen(FileName.c_str(), "wb");
		if (!f)
			Die("Cannot create %s", FileName.c_str());
		std::string out;
		for (uint32_t i = 0; i < m_DB.GetSeqCount(); ++i) {
			if ((Matched != 0) != (m_HitCounts[i] > 0))
				continue;
			std::string Label = m_DB.GetLabel(i);
			if (m_SizeOut && Matched) {
				StripAnnot(Label, "size=");
				AppendSize(Label, m_HitCounts[i]);
			}
			// the letters as the database stores them: masked by LoadDB (loaddb.cpp:117-118)
			const uint8_t *Seq = m_DB.GetSeq(i);
			if (m_Searcher) {
				m_Searcher->GetStoredTarget(i, Stored);
				Seq = (const uint8_t *)Stored.data();
			}
			AppendFasta80(out, Label.c_str(), Seq, m_DB.GetSeqLength(i));
			if (out.size() > (1u << 20)) {
				fwrite(out.data(), 1, out.size(), f);
				out.clear();
			}
		}
		fwrite(out.data(), 1, out.size(), f);
		fclose(f);
	}
}

// ------------------------------------------------------------------ OtuTabSink
static void SplitFields(const std::string &Str, std::vector<std::string> &Fields, char Sep)
{
	// myutils.cpp:1588-1607 Split with a non-zero separator: empty fields are kept, except a last one
	Fields.clear();
	std::string s;
	for (char c : Str) {
		if (c == Sep) {
			Fields.push_back(s);
			s.clear();
		} else
			s.push_back(c);
	}
	if (!s.empty())
		Fields.push_back(s);
}

static void GetStrField(const std::string &Label, const std::string &NameEq, std::string &Value)
{
	// label.cpp:27-45
	Value.clear();
	std::vector<std::string> Fields;
	SplitFields(Label, Fields, ';');
	for (const std::string &F : Fields)
		if (F.compare(0, NameEq.size(), NameEq) == 0) {
			Value = F.substr(NameEq.size());
			break;
		}
}

unsigned OtuTabSink::GetSizeFromLabel(const std::string &Label, unsigned Default)
{
	const char *p = strstr(Label.c_str(), ";size="); // label.cpp:152-161
	if (!p && Default == 0xffffffffu)
		Die("Missing size= in >%s", Label.c_str());
	return p ? (unsigned)atoi(p + 6) : Default;
}

void OtuTabSink::GetOTUNameFromLabel(const std::string &Label, std::string &OTUName)
{
	GetStrField(Label, "otu=", OTUName); // label.cpp:193-202
	if (!OTUName.empty())
		return;
	// GetAccFromLabel, label.cpp:168-182
	for (char c : Label) {
		if (c == ' ' || c == '|' || c == ';')
			if (OTUName != "gi")
				break;
		OTUName += c;
	}
	if (OTUName.empty())
		Die("Empty OTU name in label >%s", Label.c_str());
}

void OtuTabSink::GetSampleNameFromLabel(const std::string &Label, std::string &SampleName) const
{
	// label.cpp:204-234
	GetStrField(Label, "sample=", SampleName);
	if (!SampleName.empty())
		return;
	GetStrField(Label, "barcodelabel=", SampleName);
	if (!SampleName.empty())
		return;
	if (!m_SampleDelim.empty()) {
		const size_t n = Label.find(m_SampleDelim);
		if (n == std::string::npos)
			Die("delim '%s' not found in >%s", m_SampleDelim.c_str(), Label.c_str());
		SampleName = Label.substr(0, n);
		return;
	}
	for (char c : Label) {
		if (!isalpha((unsigned char)c) && !isdigit((unsigned char)c) && c != '_')
			return;
		SampleName.push_back(c);
	}
}

OtuTabSink::OtuTabSink(const std::string &OtuTabOut, const std::string &MapOut, const std::string &SampleDelim, bool Quiet,
  const std::string &BiomOut)
  : m_OtuTabOut(OtuTabOut), m_SampleDelim(SampleDelim), m_BiomOut(BiomOut), m_Quiet(Quiet)
{
	if (!MapOut.empty()) {
		m_fMap = fopen(MapOut.c_str(), "wb");
		if (!m_fMap)
			Die("Cannot create %s", MapOut.c_str());
	}
}

OtuTabSink::~OtuTabSink() { OnAllDone(); }

unsigned OtuTabSink::IndexAdd(std::vector<std::pair<std::string, unsigned>> &Map, std::vector<std::string> &Names,
  const std::string &Name, bool &Added)
{
	auto it = std::lower_bound(Map.begin(), Map.end(), Name,
	  [](const std::pair<std::string, unsigned> &a, const std::string &b) { return a.first < b; });
	Added = it == Map.end() || it->first != Name;
	if (!Added)
		return it->second;
	const unsigned Index = (unsigned)Names.size();
	Names.push_back(Name);
	Map.insert(it, std::make_pair(Name, Index));
	return Index;
}

void OtuTabSink::OnQueryDone(const SeqInfo &Query, const HitMgr &HM)
{
	const std::string QueryLabel = Query.m_Label;
	const unsigned Size = GetSizeFromLabel(QueryLabel, 1);
	m_QueryCount += Size;
	if (HM.GetHitCount() == 0)
		return;
	// HitMgr::GetTopHit (hitmgr.cpp:400-420): best float score, ties to the lowest target index
	const AlignResult *Top = &HM.m_Hits[0];
	for (const AlignResult &AR : HM.m_Hits)
		if ((float)AR.GetFractId() > (float)Top->GetFractId() ||
		    ((float)AR.GetFractId() == (float)Top->GetFractId() && AR.GetTargetIndex() < Top->GetTargetIndex()))
			Top = &AR;
	std::string OTUName, SampleName;
	GetOTUNameFromLabel(Top->GetTargetLabel(), OTUName);
	GetSampleNameFromLabel(QueryLabel, SampleName);
	m_AssignedCount += Size;
	// OTUTable::IncCount (otutab.cpp:548-556): the OTU first, then the sample
	bool added = false;
	const unsigned o = IndexAdd(m_OTUIndex, m_OTUNames, OTUName, added);
	if (added)
		m_Counts.emplace_back(m_SampleNames.size(), 0u);
	const unsigned sidx = IndexAdd(m_SampleIndex, m_SampleNames, SampleName, added);
	if (added)
		for (auto &row : m_Counts)
			row.push_back(0);
	m_Counts[o][sidx] += Size;
	if (m_fMap)
		fprintf(m_fMap, "%s\t%s\n", QueryLabel.c_str(), OTUName.c_str());
}

void OtuTabSink::OnAllDone()
{
	if (m_Done)
		return;
	m_Done = true;
	if (!m_Quiet)
		fprintf(stderr, "%u / %u mapped to OTUs (%.1f%%)\n", m_AssignedCount, m_QueryCount,
		  m_QueryCount ? 100.0 * m_AssignedCount / m_QueryCount : 0.0);
	if (m_fMap) {
		fclose(m_fMap);
		m_fMap = nullptr;
	}
	if (!m_OtuTabOut.empty()) {
		// OTUTable::ToTabbedFile (otutab.cpp:247-310)
		FILE *f = fopen(m_OtuTabOut.c_str(), "wb");
		if (!f)
			Die("Cannot create %s", m_OtuTabOut.c_str());
		fprintf(f, "#OTU ID");
		for (const std::string &sn : m_SampleNames) {
			fputc('\t', f);
			fputs(sn.c_str(), f);
		}
		fputc('\n', f);
		for (size_t o = 0; o < m_OTUNames.size(); ++o) {
			fputs(m_OTUNames[o].c_str(), f);
			for (size_t k = 0; k < m_SampleNames.size(); ++k)
				fprintf(f, "\t%u", m_Counts[o][k]);
			fputc('\n', f);
		}
		fclose(f);
	}
	if (!m_BiomOut.empty()) {
		// OTUTable::ToJsonFile (json.cpp:32-110): BIOM 1.0, sparse; an entry is followed by a comma unless it is the
		// cell of the last OTU and the last sample, whether or not that cell is written
		FILE *f = fopen(m_BiomOut.c_str(), "wb");
		if (!f)
			Die("Cannot create %s", m_BiomOut.c_str());
		const size_t NO = m_OTUNames.size(), NS = m_SampleNames.size();
		time_t now = time(nullptr);
		char when[32];
		snprintf(when, sizeof when, "%.24s", asctime(localtime(&now)));
		fprintf(f, "{\n\t\"id\":\"%s\",\n\t\"format\": \"Biological Observation Matrix 1.0\",\n", m_BiomOut.c_str());
		fprintf(f, "\t\"format_url\": \"http://biom-format.org\",\n\t\"generated_by\": \"usearch\",\n\t\"type\": \"OTU table\",\n");
		fprintf(f, "\t\"date\": \"%s\",\n\t\"matrix_type\": \"sparse\",\n\t\"matrix_element_type\": \"float\",\n", when);
		fprintf(f, "\t\"shape\": [%u,%u],\n\t\"rows\":[\n", (unsigned)NO, (unsigned)NS);
		for (size_t o = 0; o < NO; ++o)
			fprintf(f, "\t\t{\"id\":\"%s\", \"metadata\":null}%s\n", m_OTUNames[o].c_str(), o + 1 != NO ? "," : "");
		fprintf(f, "\t],\n\t\"columns\":[\n");
		for (size_t k = 0; k < NS; ++k)
			fprintf(f, "\t\t{\"id\":\"%s\", \"metadata\":null}%s\n", m_SampleNames[k].c_str(), k + 1 != NS ? "," : "");
		fprintf(f, "\t],\n\t\"data\": [\n");
		for (size_t o = 0; o < NO; ++o)
			for (size_t k = 0; k < NS; ++k)
				if (m_Counts[o][k] != 0)
					fprintf(f, "\t\t[%u,%u,%u]%s\n", (unsigned)o, (unsigned)k, m_Counts[o][k], (o + 1 < NO || k + 1 < NS) ? "," : "");
		fprintf(f, "\t]\n}\n");
		fclose(f);
	}
}

// ------------------------------------------------------------------ ClosedRefSink
ClosedRefSink::ClosedRefSink(const std::string &TabbedOut, const std::string &DbOtus, const std::string &DataOtus)
  : m_DbOtus(DbOtus), m_DataOtus(DataOtus)
{
	if (!TabbedOut.empty() && !(m_fTab = fopen(TabbedOut.c_str(), "wb")))
		Die("Cannot create %s", TabbedOut.c_str());
}

ClosedRefSink::~ClosedRefSink() { OnAllDone(); }

void ClosedRefSink::OnQueryDone(const SeqInfo &Query, const HitMgr &HM)
{
	const char *QueryLabel = Query.m_Label;
	const unsigned Size = OtuTabSink::GetSizeFromLabel(QueryLabel, 1);
	if (HM.m_Hits.empty()) {
		if (m_fTab)
			fprintf(m_fTab, "%s\t*\t*\t*\t*\t*\n", QueryLabel);
		return;
	}
	// HitMgr::GetTopHit (hitmgr.cpp:400-420): best float score, ties to the lowest target index
	const AlignResult *Top = &HM.m_Hits[0];
	for (const AlignResult &AR : HM.m_Hits)
		if ((float)AR.GetFractId() > (float)Top->GetFractId() ||
		    ((float)AR.GetFractId() == (float)Top->GetFractId() && AR.GetTargetIndex() < Top->GetTargetIndex()))
			Top = &AR;
	const unsigned TopTargetIndex = Top->GetTargetIndex();
	const double TopFractId = (float)HM.m_Hits[0].GetFractId(); // HitMgr::GetFractId(0): first in sorted order
	if (TopTargetIndex >= m_RefSeqIndexToOTUIndex.size())
		m_RefSeqIndexToOTUIndex.resize((size_t)TopTargetIndex + 1, UINT32_MAX);
	unsigned OTUIndex = m_RefSeqIndexToOTUIndex[TopTargetIndex];
	if (OTUIndex == UINT32_MAX) {
		OTUIndex = (unsigned)m_OTUTotalSize.size();
		m_RefSeqIndexToOTUIndex[TopTargetIndex] = OTUIndex;
		m_OTUTotalSize.push_back(0);
		m_OTUMemberCount.push_back(0);
		m_RefLabels.push_back(Top->GetTargetLabel());
		std::string stored;
		if (m_Searcher)
			m_Searcher->GetStoredTarget(TopTargetIndex, stored);
		m_RefSeqs.push_back(stored);
		m_DataLabels.push_back(QueryLabel);
		m_DataSeqs.push_back(std::string((const char *)Query.m_Seq, Query.m_L));
	}
	m_OTUTotalSize[OTUIndex] += Size;
	const unsigned MemberIndex = m_OTUMemberCount[OTUIndex]++;
	unsigned Ties = 0;
	std::string TiesStr;
	if (HM.m_Hits.size() > 1)
		for (const AlignResult &AR : HM.m_Hits) {
			if ((double)(float)AR.GetFractId() < TopFractId)
				break;
			if (AR.GetTargetIndex() == TopTargetIndex)
				continue;
			if (Ties > 0)
				TiesStr += ",";
			TiesStr += AR.GetTargetLabel();
			++Ties;
		}
	if (m_fTab) {
		fprintf(m_fTab, "%s\t%u\t%u\t%s\t%.1f\tties=%u", QueryLabel, OTUIndex, MemberIndex, Top->GetTargetLabel(),
		  TopFractId * 100.0, Ties);
		if (Ties > 0)
			fprintf(m_fTab, ":%s", TiesStr.c_str());
		fputc('\n', m_fTab);
	}
}

static void PsascStr(std::string &Str, const std::string &Tail) // myutils.cpp:824-839 Psasc
{
	if (!Str.empty() && Str.back() != ';')
		Str += ';';
	Str += Tail;
	if (!Str.empty() && Str.back() != ';')
		Str += ';';
}

static void WriteFasta80(FILE *f, const std::string &Label, const std::string &Seq) // seqdb.cpp:62-95 SeqToFasta
{
	if (!f || Seq.empty())
		return;
	fprintf(f, ">%s\n", Label.c_str());
	for (size_t i = 0; i < Seq.size(); i += 80) {
		fwrite(Seq.data() + i, 1, std::min<size_t>(80, Seq.size() - i), f);
		fputc('\n', f);
	}
}

// the reference's own (unstable) descending quicksort on unsigned keys (sort.h:63-102)
static void QuickSortOrderDescU(const std::vector<unsigned> &v, std::vector<unsigned> &ord, int lo, int hi)
{
	int i = lo, j = hi;
	const unsigned pivot = v[ord[(lo + hi) / 2]];
	while (i <= j) {
		while (v[ord[i]] > pivot)
			++i;
		while (v[ord[j]] < pivot)
			--j;
		if (i <= j) {
			std::swap(ord[i], ord[j]);
			++i;
			--j;
		}
	}
	if (lo < j)
		QuickSortOrderDescU(v, ord, lo, j);
	if (i < hi)
		QuickSortOrderDescU(v, ord, i, hi);
}

void ClosedRefSink::OnAllDone()
{
	if (m_Done)
		return;
	m_Done = true;
	if (m_fTab) {
		fclose(m_fTab);
		m_fTab = nullptr;
	}
	if (m_DbOtus.empty() && m_DataOtus.empty())
		return;
	const unsigned N = (unsigned)m_OTUTotalSize.size();
	std::vector<unsigned> Order(N);
	for (unsigned i = 0; i < N; ++i)
		Order[i] = i;
	if (N > 1)
		QuickSortOrderDescU(m_OTUTotalSize, Order, 0, (int)N - 1);
	FILE *fDb = m_DbOtus.empty() ? nullptr : fopen(m_DbOtus.c_str(), "wb");
	FILE *fData = m_DataOtus.empty() ? nullptr : fopen(m_DataOtus.c_str(), "wb");
	if ((!m_DbOtus.empty() && !fDb) || (!m_DataOtus.empty() && !fData))
		Die("Cannot create %s", !fDb ? m_DbOtus.c_str() : m_DataOtus.c_str());
	for (unsigned k = 0; k < N; ++k) {
		const unsigned o = Order[k];
		std::string OutRef = m_RefLabels[o], OutData = m_DataLabels[o];
		PsascStr(OutRef, "otu=" + std::to_string(k + 1) + ";size=" + std::to_string(m_OTUTotalSize[o]) + ";");
		PsascStr(OutData, "otu=" + std::to_string(k + 1) + ";ref=" + m_RefLabels[o]);
		WriteFasta80(fDb, OutRef, m_RefSeqs[o]);
		WriteFasta80(fData, OutData, m_DataSeqs[o]);
	}
	if (fDb)
		fclose(fDb);
	if (fData)
		fclose(fData);
}

// ------------------------------------------------------------------ GpuSearcher
GpuSearcher::GpuSearcher(int Device, const SeqDB &DB, const usb_params &P) : m_DB(DB), m_P(P)
{
	CheckUsb(usb_index_create(Device, &P, DB.Letters(), DB.Offsets(), DB.GetSeqCount(), &m_Index), "usb_index_create");
	CheckUsb(usb_searcher_create(m_Index, &P, &m_Searcher), "usb_searcher_create");
}

GpuSearcher::~GpuSearcher()
{
	usb_searcher_free(m_Searcher);
	usb_index_free(m_Index);
}

uint64_t GpuSearcher::GetLaunchCount() const { return usb_searcher_launch_count(m_Searcher); }

void GpuSearcher::GetStoredTarget(uint32_t Target, std::string &Seq) const
{
	const uint8_t *p = nullptr;
	uint32_t L = 0;
	CheckUsb(usb_index_seq(m_Index, Target, &p, &L), "usb_index_seq");
	Seq.assign((const char *)p, L);
}

void GpuSearcher::SetTargetAttrs(const uint32_t *LabelIds, const uint32_t *Sizes)
{
	CheckUsb(usb_index_set_attrs(m_Index, 0, m_DB.GetSeqCount(), LabelIds, Sizes), "usb_index_set_attrs");
}

// HitMgr::GetHitCount / GetHit (hitmgr.cpp:367-398,466-475): which hits of a query the sinks see with -maxhits,
// -top_hit_only, -top_hits_only.  Hits is in HitMgr::Sort order.
void SelectHits(std::vector<AlignResult> &H, const HitSelection &Sel)
{
	if (!Sel.Any() || H.empty())
		return;
	auto score = [&](const AlignResult &AR) { return AR.IsLocal() ? (float)AR.GetRawScore() : (float)AR.GetFractId(); };
	size_t n = H.size();
	if (Sel.maxhits && n > Sel.maxhits)
		n = Sel.maxhits;
	if (Sel.top_hit_only) {
		size_t top = 0; // GetTopHit over every hit (hitmgr.cpp:400-420)
		for (size_t i = 1; i < H.size(); ++i)
			if (score(H[i]) > score(H[top]) || (score(H[i]) == score(H[top]) && H[i].GetTargetIndex() < H[top].GetTargetIndex()))
				top = i;
		const AlignResult keep = H[top];
		H.assign(1, keep);
	} else {
		if (Sel.top_hits_only) {
			float best = score(H[0]);
			for (const AlignResult &AR : H)
				best = std::max(best, score(AR));
			size_t i = 1;
			while (i < n && !(score(H[i]) < best))
				++i;
			n = i;
		}
		H.resize(n);
	}
}

// Replaces the loop "SS->GetNext(Query); searcher->Search(Query)" of Thread() (search.cpp:63-86)
// for a batch: one C-ABI call, then one HitMgr per query holding AlignResults in output order.
std::shared_ptr<void> GpuSearcher::SearchRaw(const SeqDB &Queries, uint32_t First, uint32_t Count)
{
	usb_result *R = nullptr;
	if (m_QLabelIds || m_QSizes)
		CheckUsb(usb_batch_set_query_attrs(m_Searcher, Count, m_QLabelIds ? m_QLabelIds + First : nullptr,
		           m_QSizes ? m_QSizes + First : nullptr), "usb_batch_set_query_attrs");
	CheckUsb(usb_search_batch(m_Searcher, Queries.Letters(), Queries.Offsets() + First, Count, &R), "usb_search_batch");
	return std::shared_ptr<void>(R, [](void *p) { usb_result_free((usb_result *)p); });
}

void GpuSearcher::BuildHitMgrs(const std::shared_ptr<void> &Result, const SeqDB &Queries, uint32_t First, uint32_t Count,
  std::vector<HitMgr> &Out, unsigned Threads) const
{
	const usb_result *R = (const usb_result *)Result.get();
	const usb_hit *hits = usb_result_hits(R);
	const uint64_t *qoff = usb_result_query_offsets(R);
	uint64_t n_runs = 0;
	const uint32_t *runs = usb_result_runs(R, &n_runs);
	Out.resize(Count);
	auto build = [&](uint32_t q0, uint32_t q1) {
		for (uint32_t q = q0; q < q1; ++q) {
			HitMgr &HM = Out[q];
			// the paths point into the result's run arena: the first HitMgr of the batch and every
			// HitMgr with hits keep the result alive
			if (q == 0 || qoff[q] != qoff[q + 1])
				HM.m_Arena = Result;
			Queries.GetSI(First + q, HM.m_Query);
			HM.m_Hits.clear();
			for (uint64_t k = qoff[q]; k < qoff[q + 1]; ++k) {
				AlignResult AR;
				AR.m_Hit = hits[k];
				AR.m_Query = HM.m_Query;
				AR.m_Query.m_RevComp = hits[k].strand != 0;
				m_DB.GetSI(hits[k].target, AR.m_Target);
				{
					// the target as the database stores it: masked letters, like m_Target->m_Seq of the
					// reference after LoadDB (loaddb.cpp:117-118); the formats upper-case where the reference does
					const uint8_t *stored = nullptr;
					uint32_t slen = 0;
					if (usb_index_seq(m_Index, hits[k].target, &stored, &slen) == 0 && slen == AR.m_Target.m_L)
						AR.m_Target.m_Seq = stored;
				}
				AR.m_Runs = runs + hits[k].run_off;
				AR.m_Nucleo = m_P.is_nucleo != 0;
				AR.m_Local = m_P.local != 0;
				if (AR.m_Local)
					CheckUsb(usb_local_evalue(m_Searcher, hits[k].raw, hits[k].ql, &AR.m_Evalue, &AR.m_BitScore),
					  "usb_local_evalue");
				HM.m_Hits.push_back(AR);
			}
			SelectHits(HM.m_Hits, m_Sel);
		}
	};
	const unsigned T = std::max(1u, std::min(Threads, Count / 4096));
	if (T <= 1) {
		build(0, Count);
		return;
	}
	std::vector<std::thread> th;
	for (unsigned k = 0; k < T; ++k)
		th.emplace_back(build, (uint32_t)((uint64_t)Count * k / T), (uint32_t)((uint64_t)Count * (k + 1) / T));
	for (auto &t : th)
		t.join();
}

void GpuSearcher::SearchBatch(const SeqDB &Queries, uint32_t First, uint32_t Count, std::vector<HitMgr> &Out)
{
	BuildHitMgrs(SearchRaw(Queries, First, Count), Queries, First, Count, Out, 1);
}

bool GuessIsNucleo(const std::string &FastaFileName)
{
	FILE *f = fopen(FastaFileName.c_str(), "rb");
	if (!f)
		Die("Cannot open %s", FastaFileName.c_str());
	std::vector<char> buf(1u << 20);
	const size_t n = fread(buf.data(), 1, buf.size(), f);
	fclose(f);
	std::vector<char> letters;
	bool header = false;
	for (size_t i = 0; i < n; ++i) {
		const char c = buf[i];
		if (c == '>' && (i == 0 || buf[i - 1] == '\n'))
			header = true;
		else if (c == '\n')
			header = false;
		else if (!header && isalpha((unsigned char)c))
			letters.push_back(c);
	}
	if (letters.empty())
		return false;
	unsigned N = 0;
	for (unsigned k = 0; k < 100; ++k) {
		const char c = letters[(size_t)k * letters.size() / 100];
		N += strchr("ACGTUNacgtun", c) != nullptr;
	}
	return N > 80;
}

// ------------------------------------------------------------------ Search driver
uint64_t Search(const std::string &QueryFileName, const std::string &DBFileName, const SearchOpts &Opts)
{
	if (QueryFileName.empty())
		Die("Query file name not set");
	if (DBFileName.empty())
		Die("Database file name not set");
	const int ndev = usb_device_count();
	if (ndev <= 0)
		Die("No CUDA device available: this build has no CPU search path");
	const int gpus = std::min(std::max(1, Opts.gpus), ndev);
	const bool timing = getenv("USB_TIMING") != nullptr;
	auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	const double t_start = now();
	// the query file is parsed while the database is parsed, indexed and uploaded
	SeqDB DB, Q;
	std::thread parse_q([&]() { Q.FromFasta(QueryFileName); });
	usb_params P = Opts.P;
	if (IsUDBFile(DBFileName)) {
		// LoadUDB (loaddb.cpp:100-127): the stored sequences are masked already
		bool nucleo = true;
		uint32_t wl = 0;
		DB.FromUDB(DBFileName, nucleo, wl);
		if (nucleo != (P.is_nucleo != 0))
			Die("%s is a%s database, the command was set up for the other alphabet", DBFileName.c_str(), nucleo ? " nucleotide" : "n amino acid");
		P.word_length = wl;
		P.dbmask = 0;
	} else
		DB.FromFasta(DBFileName);
	const double t_db = now();
	std::vector<GpuSearcher *> searchers(gpus, nullptr);
	{
		std::vector<std::thread> th;
		for (int d = 0; d < gpus; ++d)
			th.emplace_back([&, d]() { searchers[d] = new GpuSearcher(d, DB, P); });
		for (auto &t : th)
			t.join();
	}
	parse_q.join();
	Q.DropSmallerThan(Opts.minsize);
	// Accepter rules that read labels: identities of equal labels, size= annotations
	std::vector<uint32_t> q_label, t_label, q_size, t_size;
	if (P.accept_flags & USB_ACC_NEEDS_LABELS) {
		std::unordered_map<std::string, uint32_t> ids;
		t_label.resize(DB.GetSeqCount());
		for (uint32_t i = 0; i < DB.GetSeqCount(); ++i)
			t_label[i] = ids.emplace(DB.GetLabel(i), i).first->second;
		q_label.resize(Q.GetSeqCount());
		for (uint32_t i = 0; i < Q.GetSeqCount(); ++i) {
			auto it = ids.find(Q.GetLabel(i));
			q_label[i] = it == ids.end() ? DB.GetSeqCount() + i : it->second;
		}
	}
	if (P.accept_flags & USB_ACC_NEEDS_SIZES) {
		auto size_of = [](const char *Label) { // label.cpp:152-161 with Default = UINT_MAX
			const char *p = strstr(Label, ";size=");
			if (!p)
				Die("Missing size= in >%s", Label);
			const unsigned v = (unsigned)atoi(p + 6);
			if (v == 0)
				Die("size=0 in >%s", Label);
			return v;
		};
		t_size.resize(DB.GetSeqCount());
		for (uint32_t i = 0; i < DB.GetSeqCount(); ++i)
			t_size[i] = size_of(DB.GetLabel(i));
		q_size.resize(Q.GetSeqCount());
		for (uint32_t i = 0; i < Q.GetSeqCount(); ++i)
			q_size[i] = size_of(Q.GetLabel(i));
	}
	if (Opts.ClosedRef)
		Opts.ClosedRef->SetSearcher(searchers[0]);
	for (GpuSearcher *gs : searchers) {
		gs->SetHitSelection(Opts.Sel);
		if (!t_label.empty() || !t_size.empty()) {
			gs->SetTargetAttrs(t_label.empty() ? nullptr : t_label.data(), t_size.empty() ? nullptr : t_size.data());
			gs->SetQueryAttrs(q_label.empty() ? nullptr : q_label.data(), q_size.empty() ? nullptr : q_size.data());
		}
	}
	const double t_ready = now();
	if (!Opts.quiet)
		fprintf(stderr, "%u db seqs, %u query seqs, %d GPU(s)\n", DB.GetSeqCount(), Q.GetSeqCount(), gpus);
	OutputOpts OO = Opts.Out;
	OO.nucleo = P.is_nucleo != 0;
	OO.local = P.local != 0;
	OO.match = (int)P.match;
	OO.mismatch = (int)P.mismatch;
	OutputSink Sink(OO);
	std::vector<HitSink *> ExtraSinks = Opts.ExtraSinks;
	std::unique_ptr<DBHitSink> dbhits;
	if (!Opts.dbmatched.empty() || !Opts.dbnotmatched.empty() || !Opts.dbcutout.empty()) {
		dbhits.reset(new DBHitSink(DB, Opts.dbmatched, Opts.dbnotmatched, Opts.sizein, Opts.sizeout, Opts.dbcutout, Opts.otutab));
		dbhits->SetSearcher(searchers[0]);
		ExtraSinks.push_back(dbhits.get());
	}
	const uint32_t NQ = Q.GetSeqCount();
	// an exhaustive search (-maxaccepts 0 / -maxrejects 0) keeps whole candidate lists on the device: the library
	// takes at most 2^29 (query-strand, target) pairs per batch
	uint32_t BatchSize = Opts.batch;
	if ((P.maxaccepts == 0 || P.maxrejects == 0) && DB.GetSeqCount() > 1024)
		BatchSize = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(BatchSize, (1ull << 29) / DB.GetSeqCount() / (P.strand_both ? 2 : 1)));
	const uint32_t nbatch = (NQ + BatchSize - 1) / BatchSize;
	uint64_t queries_with_hits = 0;
	// Three overlapped stages.  (1) One submitting thread per device runs usb_search_batch on the
	// batches dealt to it round-robin, at most two ahead of the consumer.  (2) The consumer takes
	// the results in input order and turns them into HitMgrs with several threads.  (3) The sinks
	// format (several threads) and write.  The output files are in query order for any GPU count
	// (the reference guarantees that only for 1 thread).
	std::vector<std::shared_ptr<void>> raw(nbatch);
	std::vector<char> ready(nbatch, 0);
	std::mutex mu;
	std::condition_variable cv;
	uint32_t consumed = 0; // batches the consumer is done with
	std::vector<std::thread> submit;
	for (int d = 0; d < gpus; ++d)
		submit.emplace_back([&, d]() {
			for (uint32_t b = (uint32_t)d; b < nbatch; b += (uint32_t)gpus) {
				{
					std::unique_lock<std::mutex> lk(mu);
					cv.wait(lk, [&]() { return b < consumed + 2u * (uint32_t)gpus; });
				}
				const uint32_t first = b * BatchSize;
				std::shared_ptr<void> r = searchers[d]->SearchRaw(Q, first, std::min<uint32_t>(BatchSize, NQ - first));
				{
					std::lock_guard<std::mutex> lk(mu);
					raw[b] = std::move(r);
					ready[b] = 1;
				}
				cv.notify_all();
			}
		});
	const unsigned host_threads = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
	double t_wait = 0, t_sink = 0, t_build = 0;
	std::vector<HitMgr> batch;
	for (uint32_t b = 0; b < nbatch; ++b) {
		const double t0 = now();
		std::shared_ptr<void> r;
		{
			std::unique_lock<std::mutex> lk(mu);
			cv.wait(lk, [&]() { return ready[b] != 0; });
			r = std::move(raw[b]);
		}
		const double t1 = now();
		const uint32_t first = b * BatchSize, count = std::min<uint32_t>(BatchSize, NQ - first);
		searchers[b % gpus]->BuildHitMgrs(r, Q, first, count, batch, host_threads);
		r.reset();
		const double t2 = now();
		Sink.OnBatchDone(batch);
		for (HitSink *x : ExtraSinks)
			x->OnBatchDone(batch);
		for (const HitMgr &HM : batch)
			queries_with_hits += HM.GetHitCount() > 0;
		batch.clear();
		{
			std::lock_guard<std::mutex> lk(mu);
			consumed = b + 1;
		}
		cv.notify_all();
		const double t3 = now();
		t_wait += t1 - t0;
		t_build += t2 - t1;
		t_sink += t3 - t2;
	}
	for (auto &t : submit)
		t.join();
	Sink.OnAllDone();
	for (HitSink *x : ExtraSinks)
		x->OnAllDone();
	for (GpuSearcher *s : searchers)
		delete s;
	if (timing)
		fprintf(stderr, "timing: db parse %.2fs, index+upload (query parse alongside) %.2fs, search+output %.2fs (hit lists %.2fs, sinks %.2fs, waiting for the GPU %.2fs)\n",
		  t_db - t_start, t_ready - t_db, now() - t_ready, t_build, t_sink, t_wait);
	return queries_with_hits;
}

} // namespace usbhost
