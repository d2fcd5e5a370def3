// usb_host.h -- host-side mirror of the reference's search framework for the GPU hot path.
//
// Same names, argument meaning and error behaviour as the reference classes they stand in for
// (SURVEY.md section 8b), re-implemented over the C ABI of include/usb200.h:
//   SeqInfo / SeqDB            seqinfo.h, seqdb.cpp:256 (GetSI), fastaseqsource.cpp:25-124
//   AlignResult                alignresult.h:17-245, arscorer.cpp:201-296 (FillLo)
//   HitMgr                     hitmgr.cpp:120-161,400,477
//   HitSink / OutputSink       hitsink.h:57, outputsink.cpp:358, outputuc.cpp, blast6out.cpp, userout.cpp
//   Searcher / GpuSearcher     searcher.h:21-96 -- Search(Query) becomes SearchBatch(first,n)
//   Search()                   search.cpp:89 (driver: load DB, fan out, sinks in input order)
// Error convention: Die() prints to stderr and exits 1 (myutils.cpp:867).
#pragma once
#include <cstdint>
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

#include "../../../include/usb200.h"

namespace usbhost {

[[noreturn]] void Die(const char *fmt, ...);
void Warning(const char *fmt, ...);
void StripAnnot(std::string &Label, const std::string &NameEq); // label.cpp:47-75
void AppendSize(std::string &Label, unsigned Size);             // label.cpp:88-91

struct SeqInfo {
	const char *m_Label = nullptr;
	const uint8_t *m_Seq = nullptr;
	uint32_t m_L = 0;
	uint32_t m_Index = 0;
	bool m_RevComp = false;
	const char *m_Qual = nullptr; // FASTQ input only: quality characters, one per letter
};

// In-memory sequence set (seqdb.h); letters are kept exactly as read.
class SeqDB {
public:
	// FASTA, or FASTQ when the file starts with '@' (MakeSeqSource seqsource.cpp:37-61, GetFileType filetype.cpp:7-47)
	void FromFasta(const std::string &FileName);
	// SeqDB part of a .udb file (UDBData::FromUDBFile, udbio.cpp:242-279): letters come back masked
	// as stored; returns the alphabet and word width of the file's header
	void FromUDB(const std::string &FileName, bool &IsNucleo, uint32_t &WordLength);
	uint32_t GetSeqCount() const { return (uint32_t)m_Offsets.size() - 1; }
	void GetSI(uint32_t Index, SeqInfo &SI) const;
	// Drops the sequences whose size= annotation is below MinSize (the query loop of Thread(), search.cpp:59-82,
	// skips them before the search: they reach no sink); a label without size= is an error (label.cpp:152-161)
	void DropSmallerThan(unsigned MinSize);
	const uint8_t *GetSeq(uint32_t i) const { return m_Letters.data() + m_Offsets[i]; }
	uint32_t GetSeqLength(uint32_t i) const { return (uint32_t)(m_Offsets[i + 1] - m_Offsets[i]); }
	const char *GetLabel(uint32_t i) const { return m_Labels[i].c_str(); }
	const uint8_t *Letters() const { return m_Letters.data(); }
	const uint64_t *Offsets() const { return m_Offsets.data(); }

private:
	void FromFastq(const char *p, const char *end, const std::string &FileName);
	std::vector<uint8_t> m_Letters;
	std::vector<char> m_Quals; // FASTQ input: same offsets as the letters
	bool m_HasQual = false;
	std::vector<uint64_t> m_Offsets{0};
	std::vector<std::string> m_Labels;
};

// One accepted alignment: the statistics the reference derives lazily from (Query, Target, Path).
class AlignResult {
public:
	SeqInfo m_Query, m_Target;
	usb_hit m_Hit;
	const uint32_t *m_Runs = nullptr; // (length << 2) | op, op 0=M 1=D 2=I

	double GetFractId() const { return m_Hit.alnlen == 0 ? 0.0 : double(m_Hit.ids) / double(m_Hit.alnlen); }
	double GetPctId() const { return 100.0 * GetFractId(); }
	const char *GetQueryLabel() const { return m_Query.m_Label; }
	const char *GetTargetLabel() const { return m_Target.m_Label; }
	unsigned GetTargetIndex() const { return m_Hit.target; }
	unsigned GetIQL() const { return m_Hit.ql; }
	unsigned GetITL() const { return m_Hit.tl; }
	unsigned GetAlnLength() const { return m_Hit.alnlen; }
	unsigned GetIdCount() const { return m_Hit.ids; }
	unsigned GetMismatchCount() const { return m_Hit.mism; }
	unsigned GetGapCount() const { return m_Hit.intgaps; }
	unsigned GetGapOpenCount() const { return m_Hit.opens; }
	unsigned GetGapExtCount() const { return m_Hit.intgaps - m_Hit.opens; }
	unsigned GetLetterPairCount() const { return m_Hit.ids + m_Hit.mism; }
	unsigned GetDiffCount() const { return m_Hit.mism + m_Hit.intgaps; }
	unsigned GetPathLength() const;
	unsigned GetAllGapCount() const { return m_Hit.intgaps + (GetPathLength() - m_Hit.alnlen); }
	// m_HSP: global alignments span both sequences (alignresult.cpp:137-145); local ones carry
	// the aligned segment (alignresult.cpp:173), which starts and ends with an M column
	bool m_Local = false, m_Nucleo = true;
	double m_Evalue = -1.0, m_BitScore = 0.0; // estats.cpp:73-96, filled by the searcher for local hits
	unsigned GetLoi() const { return m_Local ? m_Hit.first_mq : 0; }
	unsigned GetHii() const { return m_Local ? m_Hit.last_mq : m_Hit.ql - 1; }
	unsigned GetLoj() const { return m_Local ? m_Hit.first_mt : 0; }
	unsigned GetHij() const { return m_Local ? m_Hit.last_mt : m_Hit.tl - 1; }
	// arscorer.cpp:683-745 (no ORFs on this path): query coordinates are on the plus strand
	unsigned GetIQLo() const { return m_Hit.strand ? m_Hit.ql - GetHii() - 1 : GetLoi(); }
	unsigned GetIQHi() const { return m_Hit.strand ? m_Hit.ql - GetLoi() - 1 : GetHii(); }
	unsigned GetITLo() const { return GetLoj(); }
	unsigned GetITHi() const { return GetHij(); }
	unsigned GetIQLo1() const { return GetIQLo() + 1; }
	unsigned GetIQHi1() const { return GetIQHi() + 1; }
	unsigned GetITLo1() const { return GetITLo() + 1; }
	unsigned GetITHi1() const { return GetITHi() + 1; }
	unsigned GetTLo6() const { return m_Hit.strand ? GetITHi1() : GetITLo1(); } // arscorer.cpp:748-808
	unsigned GetTHi6() const { return m_Hit.strand ? GetITLo1() : GetITHi1(); }
	double GetRawScore() const { return m_Local ? (double)m_Hit.raw : 0.0; }   // arscorer.cpp:87-103
	double GetEvalue() const { return m_Local ? m_Evalue : -1.0; }             // arscorer.cpp:69-85
	double GetBitScore() const { return m_Local ? m_BitScore : 0.0; }          // arscorer.cpp:105-120
	bool IsLocal() const { return m_Local; }
	unsigned GetQLoT() const { return m_Hit.first_mq; }
	unsigned GetQHiT() const { return m_Hit.last_mq; }
	unsigned GetTLoT() const { return m_Hit.first_mt; }
	unsigned GetTHiT() const { return m_Hit.last_mt; }
	unsigned GetQUnT() const { return m_Hit.ql - m_Hit.last_mq - 1; }
	unsigned GetTUnT() const { return m_Hit.tl - m_Hit.last_mt - 1; }
	char GetQueryStrand() const { return !m_Nucleo ? '.' : m_Hit.strand ? '-' : '+'; } // arscorer.cpp:156-176
	char GetTargetStrand() const { return m_Nucleo ? '+' : '.'; }
	// segment lengths m_HSP.Leni / Lenj: the whole sequences for a global alignment (alignresult.cpp:137-145)
	unsigned GetQuerySegLength() const { return m_Local ? m_Hit.last_mq - m_Hit.first_mq + 1 : m_Hit.ql; }
	unsigned GetTargetSegLength() const { return m_Local ? m_Hit.last_mt - m_Hit.first_mt + 1 : m_Hit.tl; }
	unsigned GetTermGapCount() const { return GetPathLength() - m_Hit.alnlen; }
	unsigned GetEditDiffCount() const { return m_Hit.mism + m_Hit.intgaps + GetTermGapCount(); } // alignresult.h:160
	double GetFractMatchId() const { return m_Hit.ids == 0 ? 0.0 : double(m_Hit.ids) / double(m_Hit.ids + m_Hit.mism); }
	double GetPctGaps() const { return m_Hit.alnlen == 0 ? 0.0 : 100.0 * (double(m_Hit.intgaps) / double(m_Hit.alnlen)); }
	double GetQueryCov() const // arscorer.cpp:122-137
	{
		return m_Local ? double(GetQuerySegLength()) / double(m_Hit.ql) : double(m_Hit.last_mq - m_Hit.first_mq + 1) / m_Hit.ql;
	}
	double GetTargetCov() const // arscorer.cpp:139-154
	{
		return m_Local ? double(GetTargetSegLength()) / double(m_Hit.tl) : double(m_Hit.ids + m_Hit.mism) / double(m_Hit.tl);
	}
	// letters of the query as aligned: reverse-complemented for a minus-strand hit (the reference aligns a
	// second SeqInfo made by SeqInfo::GetRevComp, seqinfo.cpp:292-325); Buf is used only in that case
	const uint8_t *GetQSeq(std::string &Buf) const;
	void GetPath(std::string &Path) const;           // pathinfo.cpp:37-214
	void GetCompressedPath(std::string &CPath) const; // comppath.cpp:7-48
};

// Per-query hit list in output order (hitmgr.cpp).
class HitMgr {
public:
	SeqInfo m_Query;
	std::vector<AlignResult> m_Hits;
	std::shared_ptr<void> m_Arena; // keeps the batch result (run arena the hits' paths point into) alive
	unsigned GetHitCount() const { return (unsigned)m_Hits.size(); }
	const AlignResult *GetTopHit() const { return m_Hits.empty() ? nullptr : &m_Hits[0]; }
};

class HitSink {
public:
	virtual ~HitSink() {}
	virtual void OnQueryDone(const SeqInfo &Query, const HitMgr &HM) = 0;
	// all queries of a batch, in input order (default: one OnQueryDone per query)
	virtual void OnBatchDone(const std::vector<HitMgr> &Batch)
	{
		for (const HitMgr &HM : Batch)
			OnQueryDone(HM.m_Query, HM);
	}
	virtual void OnAllDone() {}
};

class GpuSearcher;

struct OutputOpts {
	std::string uc, blast6out, userout, userfields;
	// outputsink.cpp:135-195 OpenOutputFiles: the other per-hit and per-query files
	std::string alnout, fastapairs, qsegout, tsegout, matched, notmatched, matchedfq, notmatchedfq, trimout;
	std::string cmdline;           // first line of -alnout (PrintCmdLine, myutils.cpp:1667)
	unsigned rowlen = 80;          // -rowlen (o_defaults.inc:53)
	unsigned flank = 8;            // -flank (o_defaults.inc:39), userfield qsegf
	unsigned wordlength = 8;       // -wordlength as the userfield kmerid reads it (arscorer.cpp:886)
	bool output_no_hits = false;
	bool uc_hitsonly = false;      // -uc_hitsonly: no N records (outputuc.cpp:14-15)
	// alphabet and substitution scores behind the annotation row and the positives count
	// (g_SubstMx: setnucmx.cpp:33-87 with -match/-mismatch, or BLOSUM62 blosum62.cpp:17-96)
	bool nucleo = true, local = false;
	int match = 1, mismatch = -2;
};

// HitMgr::GetHitCount / GetHit (hitmgr.cpp:367-398,466-475): which of a query's hits the sinks see
struct HitSelection {
	unsigned maxhits = 0;      // -maxhits, 0 = not set
	bool top_hit_only = false; // -top_hit_only: GetTopHit (best score, ties to the lowest target index)
	bool top_hits_only = false; // -top_hits_only: the hits that share the best score
	bool Any() const { return maxhits != 0 || top_hit_only || top_hits_only; }
};

// applies a HitSelection to a query's hit list (used by GpuSearcher::BuildHitMgrs)
void SelectHits(std::vector<AlignResult> &Hits, const HitSelection &Sel);

struct FormatTables; // identity / substitution tables of the row formats (usb_host.cpp)

// outputsink.cpp:358-381; formats of outputuc.cpp:19-69, blast6out.cpp:27-80, userout.cpp:126-352,
// alnout.cpp:45-171 (+ the per-query report outputsink.cpp:237-356), fastapairs / qsegout / tsegout
// outputsink.cpp:17-44,197-235, matched / notmatched outputsink.cpp:383-400
class OutputSink : public HitSink {
public:
	enum Stream { O_UC, O_B6, O_USER, O_ALN, O_PAIRS, O_QSEG, O_TSEG, O_MATCHED, O_NOTMATCHED, O_MATCHEDFQ, O_NOTMATCHEDFQ, O_TRIM, O_COUNT };
	explicit OutputSink(const OutputOpts &O);
	~OutputSink() override;
	void OnQueryDone(const SeqInfo &Query, const HitMgr &HM) override;
	// the reference formats under one lock (outputsink.cpp:360,380); here the queries of a batch
	// are formatted by several threads into per-chunk buffers that are written in input order
	void OnBatchDone(const std::vector<HitMgr> &Batch) override;
	void OnAllDone() override;

private:
	typedef std::string Bufs[O_COUNT];
	void FormatQuery(const SeqInfo &Query, const HitMgr &HM, Bufs &out) const;
	void Flush(int k, std::string &buf, bool force);
	void OutputUC(const SeqInfo &Query, const HitMgr &HM, std::string &out) const;
	void OutputBlast6(const HitMgr &HM, std::string &out) const;
	void OutputUser(const HitMgr &HM, std::string &out) const;
	void OutputReport(const SeqInfo &Query, const HitMgr &HM, std::string &out) const;
	void OutputAln(const AlignResult &AR, std::string &out) const;
	FILE *m_f[O_COUNT] = {};
	Bufs m_b;
	OutputOpts m_O;
	bool m_OutputNoHits = false;
	std::vector<int> m_UserFields;
	std::shared_ptr<FormatTables> m_T;
};

// -dbmatched / -dbnotmatched (dbhitsink.cpp:42-159): the database sequences with / without hits, in
// database order, written when the search is over.  The letters are the ones the database stores
// (masked); with -sizeout the matched labels get the number of hits (query size= with -sizein).
class DBHitSink : public HitSink {
public:
	DBHitSink(const SeqDB &DB, const std::string &DbMatched, const std::string &DbNotMatched, bool SizeIn, bool SizeOut,
	  const std::string &DbCutOut = std::string(), // -dbcutout: the hit segments (dbhitsink.cpp:52-106)
	  bool TopHitOnly = false);                    // -otutab counts only a query's first hit (dbhitsink.cpp:138-139)
	~DBHitSink() override;
	void SetSearcher(const GpuSearcher *S) { m_Searcher = S; }
	void OnQueryDone(const SeqInfo &Query, const HitMgr &HM) override;
	void OnAllDone() override;

private:
	const SeqDB &m_DB;
	const GpuSearcher *m_Searcher = nullptr;
	std::string m_DbMatched, m_DbNotMatched, m_DbCutOut;
	std::vector<std::vector<unsigned>> m_Los, m_His; // -dbcutout: first / last aligned target position of every hit
	bool m_SizeIn, m_SizeOut, m_TopHitOnly = false, m_Done = false;
	std::vector<unsigned> m_HitCounts;
};

// OTU table of -otutab (otutabsink.cpp:25-76, otutab.cpp:247-310,444-556, label.cpp:152-234): every
// query adds its size= annotation (default 1) to the cell (OTU of its top hit, its sample).  OTUs
// and samples are numbered in order of first appearance, which with input-order draining is the
// reference's order for -threads 1.
class OtuTabSink : public HitSink {
public:
	OtuTabSink(const std::string &OtuTabOut, const std::string &MapOut, const std::string &SampleDelim, bool Quiet,
	  const std::string &BiomOut = std::string()); // -biomout: OTUTable::ToJsonFile (json.cpp:32-110)
	~OtuTabSink() override;
	void OnQueryDone(const SeqInfo &Query, const HitMgr &HM) override;
	void OnAllDone() override;
	static unsigned GetSizeFromLabel(const std::string &Label, unsigned Default);
	static void GetOTUNameFromLabel(const std::string &Label, std::string &OTUName);
	void GetSampleNameFromLabel(const std::string &Label, std::string &SampleName) const;

private:
	std::string m_OtuTabOut, m_SampleDelim, m_BiomOut;
	FILE *m_fMap = nullptr;
	bool m_Quiet = false, m_Done = false;
	std::vector<std::string> m_OTUNames, m_SampleNames;
	std::vector<std::vector<unsigned>> m_Counts; // [otu][sample]
	std::vector<std::pair<std::string, unsigned>> m_OTUIndex, m_SampleIndex; // sorted lookup tables
	unsigned m_AssignedCount = 0, m_QueryCount = 0;
	unsigned IndexAdd(std::vector<std::pair<std::string, unsigned>> &Map, std::vector<std::string> &Names, const std::string &Name,
	  bool &Added);
};

class GpuSearcher;

// -closed_ref (closedrefsink.cpp:34-165): every query goes to the reference sequence of its top hit;
// the reference sequences that were hit become OTUs numbered in order of first use.  -tabbedout gets
// one line per query, -dbotus / -dataotus the OTUs' reference sequences (as stored in the database,
// i.e. masked) and first member reads in decreasing size order.
class ClosedRefSink : public HitSink {
public:
	ClosedRefSink(const std::string &TabbedOut, const std::string &DbOtus, const std::string &DataOtus);
	~ClosedRefSink() override;
	void SetSearcher(const GpuSearcher *S) { m_Searcher = S; }
	void OnQueryDone(const SeqInfo &Query, const HitMgr &HM) override;
	void OnAllDone() override;

private:
	const GpuSearcher *m_Searcher = nullptr;
	FILE *m_fTab = nullptr;
	std::string m_DbOtus, m_DataOtus;
	bool m_Done = false;
	std::vector<unsigned> m_RefSeqIndexToOTUIndex, m_OTUTotalSize, m_OTUMemberCount;
	std::vector<std::string> m_RefLabels, m_RefSeqs, m_DataLabels, m_DataSeqs;
};

// Abstract searcher (searcher.h:21-96), batched.
class Searcher {
public:
	virtual ~Searcher() {}
	// Searches queries [First, First+Count) of Queries and appends one HitMgr per query to Out.
	virtual void SearchBatch(const SeqDB &Queries, uint32_t First, uint32_t Count, std::vector<HitMgr> &Out) = 0;
};

// UDBUsortedSearcher + GlobalAligner + Accepter + Terminator on one GPU.
class GpuSearcher : public Searcher {
public:
	GpuSearcher(int Device, const SeqDB &DB, const usb_params &P);
	~GpuSearcher() override;
	void SearchBatch(const SeqDB &Queries, uint32_t First, uint32_t Count, std::vector<HitMgr> &Out) override;
	// the two halves of SearchBatch, for callers that overlap them: the C-ABI call (one thread per
	// device) and the construction of the HitMgrs from its result (any number of threads)
	std::shared_ptr<void> SearchRaw(const SeqDB &Queries, uint32_t First, uint32_t Count);
	void BuildHitMgrs(const std::shared_ptr<void> &Result, const SeqDB &Queries, uint32_t First, uint32_t Count,
	  std::vector<HitMgr> &Out, unsigned Threads) const;
	uint64_t GetLaunchCount() const;
	void SetHitSelection(const HitSelection &Sel) { m_Sel = Sel; }
	// label identities / size= annotations of the queries (indexed like the query SeqDB), for the
	// Accepter rules that read labels; the pointers must stay valid while batches are searched
	void SetQueryAttrs(const uint32_t *LabelIds, const uint32_t *Sizes)
	{
		m_QLabelIds = LabelIds;
		m_QSizes = Sizes;
	}
	void SetTargetAttrs(const uint32_t *LabelIds, const uint32_t *Sizes);
	// letters of a target as the index stores them (masked), usb_index_seq
	void GetStoredTarget(uint32_t Target, std::string &Seq) const;

private:
	HitSelection m_Sel;
	const uint32_t *m_QLabelIds = nullptr, *m_QSizes = nullptr;
	const SeqDB &m_DB;
	usb_params m_P;
	usb_index *m_Index = nullptr;
	usb_searcher *m_Searcher = nullptr;
};

struct SearchOpts {
	usb_params P;
	OutputOpts Out;
	std::vector<HitSink *> ExtraSinks; // run after the OutputSink for every batch (not owned)
	ClosedRefSink *ClosedRef = nullptr; // one of ExtraSinks: gets the searcher for the stored target letters
	std::string dbmatched, dbnotmatched, dbcutout; // -dbmatched / -dbnotmatched / -dbcutout: Search() adds a DBHitSink
	bool sizein = false, sizeout = false; // -sizein / -sizeout as -dbmatched reads them
	bool otutab = false;                  // cmd_otutab: the DBHitSink counts one hit per query
	unsigned minsize = 0;                 // -minsize: queries with a smaller size= annotation are not searched (search.cpp:59-82)
	int gpus = 1;
	uint32_t batch = 1u << 18;
	bool quiet = false;
	HitSelection Sel;
};

struct ClusterOpts {
	usb_params P;             // usb_default_params(&P, 1) + -id
	std::string uc, centroids, sort, relabel;
	// the per-hit files of the OutputSink that MakeClusterSearcher puts behind the searcher next to the ClusterSink
	// (-userout, -blast6out, -alnout, -fastapairs, -qsegout, -tsegout, -matched, -notmatched); Out.uc stays empty, the
	// .uc file is written by the cluster loop itself
	OutputOpts Out;
	uint32_t max_block = 1u << 16;
	unsigned minsize = 0;     // -minsize: centroids of smaller clusters are not written
	bool minsize_filled = false;
	bool sizein = false, sizeout = false;
	bool quiet = false;
	// -cluster_smallmem (clustersmallmem.cpp:50-143): no dereplication, input order, the input must be
	// sorted as -sortedby says (length by default; size; other = no check)
	bool smallmem = false;
	std::string sortedby;
};

// clusterfast.cpp:81 ClusterFast() with -threads 1 semantics (or ClusterSmallmem, clustersmallmem.cpp:50,
// when Opts.smallmem); returns the number of clusters.
uint64_t ClusterFast(const std::string &ReadsFileName, const ClusterOpts &Opts);

struct UniquesOpts {
	std::string fastaout, relabel;
	bool sizeout = false;
	bool sizein = false;       // sizes = sums of the members' size= annotations (derepresult.cpp:211-225,822-844)
	unsigned minuniquesize = 0;
	unsigned topn = 0;         // -topn: at most this many uniques in -fastaout (derepresult.cpp:705-707), 0 = all
};
// DerepResult::Write (derepresult.cpp:878-895) for a finished grouping: UniqOf[i] = unique of input sequence i,
// uniques numbered in order of first occurrence.  Host only.
void WriteUniques(const SeqDB &Input, const std::vector<unsigned> &UniqOf, unsigned UniqueCount, const UniquesOpts &Opts);
// -fastx_uniques (derepfull.cpp:214-236): full-length dereplication on the device (usb_derep_full), output on the
// host; returns the number of uniques.
uint64_t FastxUniques(const std::string &InputFileName, const UniquesOpts &Opts);

// SeqDB::GetIsNucleo (seqdb.cpp:268-310): a database is nucleotide when more than 80 of 100 sampled
// letters are ACGTUN (either case).  The reference samples with rand(); here the sample is an even
// stride over the first sequences of the file (deterministic).
bool GuessIsNucleo(const std::string &FastaFileName);

// makeudb.cpp:27-60 cmd_makeudb_usearch: mask, index, write a .udb the reference can load too.
void MakeUDB(const std::string &FastaFileName, const std::string &OutputFileName, const usb_params &P);

// True when the file starts with the .udb magic (loaddb.cpp:100-107 chooses the loader the same way).
bool IsUDBFile(const std::string &FileName);
// Alphabet recorded in a .udb header.
bool UDBIsNucleo(const std::string &FileName);

// search.cpp:89 Search(): returns the number of queries with at least one hit.  DBFileName may be
// a FASTA file (masked and indexed here, LoadDB loaddb.cpp:129) or a .udb file.
uint64_t Search(const std::string &QueryFileName, const std::string &DBFileName, const SearchOpts &Opts);

} // namespace usbhost
