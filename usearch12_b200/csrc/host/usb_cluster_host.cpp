// usb_cluster_host.cpp -- host driver of -cluster_fast over usb_cluster_round.
//
// Mirrors ClusterFast() (clusterfast.cpp:81-133) with -threads 1 semantics: dereplication
// (derepfull.cpp:130-212; equality is case-insensitive, uniques in first-occurrence order),
// optional -sort length|size (clusterfast.cpp:38-79, the reference's own quicksort sort.h:63-102),
// the greedy centroid loop (batched: see usb_cluster.inc), ClusterSink bookkeeping
// (clustersink.cpp:306-359), .uc S/H/C records including the dereplicated members
// (outputuc.cpp:9-92, clustersink.cpp:477-492) and the centroids FASTA in decreasing cluster
// size order, 80 letters per line (clustersink.cpp:246-272).
#include <algorithm>
#include <cctype>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "usb_host.h"

#include <memory>

#include <chrono>
#include <climits>
#include <thread>
#include <cstring>

namespace usbhost {

static void CheckUsb2(int rc, const char *what)
{
	if (rc != 0)
		Die("%s: %s (usb200 error %d)", what, usb_last_error(), rc);
}

// sort.h:63-102,132 QuickSortOrderDesc<unsigned>: middle pivot, Hoare partition, not stable
static void QuickSortOrderDescRecurse(const unsigned *Values, int left, int right, unsigned *Order)
{
	int i = left, j = right;
	const unsigned pivot = Values[Order[(left + right) / 2]];
	while (i <= j) {
		while (Values[Order[i]] > pivot)
			++i;
		while (Values[Order[j]] < pivot)
			--j;
		if (i <= j) {
			std::swap(Order[i], Order[j]);
			++i;
			--j;
		}
	}
	if (left < j)
		QuickSortOrderDescRecurse(Values, left, j, Order);
	if (i < right)
		QuickSortOrderDescRecurse(Values, i, right, Order);
}

static void QuickSortOrderDesc(const std::vector<unsigned> &Values, std::vector<unsigned> &Order)
{
	Order.resize(Values.size());
	for (unsigned i = 0; i < Order.size(); ++i)
		Order[i] = i;
	if (!Values.empty())
		QuickSortOrderDescRecurse(Values.data(), 0, (int)Values.size() - 1, Order.data());
}

static void appendf2(std::string &s, const char *fmt, ...)
{
	char tmp[256];
	va_list ap;
	va_start(ap, fmt);
	int n = vsnprintf(tmp, sizeof tmp, fmt, ap);
	va_end(ap);
	if (n > 0)
		s.append(tmp, (size_t)std::min<int>(n, (int)sizeof tmp - 1));
}

// DerepFull (derepfull.cpp:130-212) with -threads 1 semantics: case-insensitive equality, uniques in
// first-occurrence order.  UniqOf[i] = unique of sequence i, First[u] = its first member, USize[u].
// The grouping runs on the device (usb_derep_full); the letters are flattened for it.
static void DerepFullDevice(const SeqDB &Input, std::vector<unsigned> &UniqOf, std::vector<unsigned> &First,
  std::vector<unsigned> &USize)
{
	const unsigned SeqCount = Input.GetSeqCount();
	std::vector<uint64_t> Off((size_t)SeqCount + 1, 0);
	for (unsigned i = 0; i < SeqCount; ++i)
		Off[i + 1] = Off[i] + Input.GetSeqLength(i);
	std::vector<uint8_t> Letters(Off[SeqCount] + 16);
	{
		const unsigned T = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
		std::vector<std::thread> th;
		for (unsigned k = 0; k < T; ++k)
			th.emplace_back([&, k]() {
				const unsigned a = (unsigned)((uint64_t)SeqCount * k / T), b = (unsigned)((uint64_t)SeqCount * (k + 1) / T);
				for (unsigned i = a; i < b; ++i)
					memcpy(Letters.data() + Off[i], Input.GetSeq(i), Input.GetSeqLength(i));
			});
		for (auto &t : th)
			t.join();
	}
	UniqOf.assign(SeqCount, 0);
	uint32_t nu = 0;
	CheckUsb2(usb_derep_full(0, Letters.data(), Off.data(), SeqCount, UniqOf.data(), &nu), "usb_derep_full");
	First.assign(nu, UINT_MAX);
	USize.assign(nu, 0);
	for (unsigned i = 0; i < SeqCount; ++i) {
		const unsigned u = UniqOf[i];
		if (First[u] == UINT_MAX)
			First[u] = i;
		++USize[u];
	}
}

// label.cpp:47-75 StripAnnot
void StripAnnot(std::string &Label, const std::string &NameEq)
{
	if (Label.find(NameEq) == std::string::npos)
		return;
	std::string NewLabel, f;
	std::vector<std::string> Fields;
	for (char c : Label) { // myutils.cpp:1588-1607 Split(';')
		if (c == ';') {
			Fields.push_back(f);
			f.clear();
		} else
			f.push_back(c);
	}
	if (!f.empty())
		Fields.push_back(f);
	for (const std::string &F : Fields) {
		if (F.find(NameEq) == 0)
			continue;
		NewLabel += F + ";";
	}
	if (NewLabel.find('=') == std::string::npos)
		Label = NewLabel.empty() ? NewLabel : NewLabel.substr(0, NewLabel.size() - 1);
	else
		Label = NewLabel;
}

// label.cpp:152-161 GetSizeFromLabel
static unsigned GetSizeFromLabel(const char *Label, unsigned Default)
{
	const char *p = strstr(Label, ";size=");
	if (p)
		return (unsigned)atoi(p + 6);
	if (Default == UINT_MAX)
		Die("Missing size= in >%s", Label);
	return Default;
}

// label.cpp:88-91 AppendSize -> AppendIntField (myutils.cpp:824-839 Psasc)
void AppendSize(std::string &Label, unsigned Size)
{
	if (!Label.empty() && Label.back() != ';')
		Label += ';';
	Label += "size=" + std::to_string(Size) + ";";
}

// DerepResult::Write (derepresult.cpp:878-895): -fastaout (ToFastx :689-775: uniques in order of decreasing size --
// the reference's own quicksort, not stable --, labelled with the first member's label, optionally relabelled and
// annotated with ;size=N;).  -uc and -tabbedout are not offered: the reference binary stops with an assert before it
// writes them (progress.cpp:496, nested ProgressStartOther in DerepResult::Write), so there is nothing to match.
void WriteUniques(const SeqDB &Input, const std::vector<unsigned> &UniqOf, unsigned UniqueCount, const UniquesOpts &Opts)
{
	const unsigned SeqCount = Input.GetSeqCount();
	// members of every unique in input order (DerepResult::GetSeqIndex)
	std::vector<unsigned> MemberOff(UniqueCount + 1, 0), Members(SeqCount);
	for (unsigned i = 0; i < SeqCount; ++i)
		++MemberOff[UniqOf[i] + 1];
	for (unsigned u = 0; u < UniqueCount; ++u)
		MemberOff[u + 1] += MemberOff[u];
	{
		std::vector<unsigned> cur(MemberOff.begin(), MemberOff.end() - 1);
		for (unsigned i = 0; i < SeqCount; ++i)
			Members[cur[UniqOf[i]]++] = i;
	}
	// SetSizes (derepresult.cpp:822-844): member counts, or the sums of the size= annotations with -sizein
	std::vector<unsigned> Sizes(UniqueCount), Order;
	for (unsigned u = 0; u < UniqueCount; ++u) {
		unsigned Size = MemberOff[u + 1] - MemberOff[u];
		if (Opts.sizein) {
			Size = 0;
			for (unsigned m = MemberOff[u]; m < MemberOff[u + 1]; ++m)
				Size += GetSizeFromLabel(Input.GetLabel(Members[m]), 1);
		}
		Sizes[u] = Size;
	}
	QuickSortOrderDesc(Sizes, Order);
	auto open = [](const std::string &fn) {
		FILE *f = fopen(fn.c_str(), "wb");
		if (!f)
			Die("Cannot create %s", fn.c_str());
		return f;
	};
	auto flush = [](FILE *f, std::string &out, bool force) {
		if (force || out.size() > (1u << 20)) {
			fwrite(out.data(), 1, out.size(), f);
			out.clear();
		}
	};
	std::string out;
	if (!Opts.fastaout.empty()) {
		FILE *f = open(Opts.fastaout);
		unsigned N = UniqueCount, counter = 0;
		if (Opts.topn != 0 && N > Opts.topn)
			N = Opts.topn;
		for (unsigned k = 0; k < N; ++k) {
			const unsigned u = Order[k], Size = Sizes[u], r = Members[MemberOff[u]];
			if (Size < Opts.minuniquesize)
				break;
			std::string Label = Input.GetLabel(r); // DerepResult::MakeLabel (derepresult.cpp:255-284)
			if (!Opts.relabel.empty())
				Label = Opts.relabel + std::to_string(++counter);
			if (Opts.sizeout) {
				StripAnnot(Label, "size=");
				AppendSize(Label, Size);
			}
			const uint8_t *s = Input.GetSeq(r);
			const unsigned L = Input.GetSeqLength(r);
			out += '>';
			out += Label;
			out += '\n';
			for (unsigned i = 0; i < L; i += 80) {
				out.append((const char *)s + i, std::min(80u, L - i));
				out += '\n';
			}
			flush(f, out, false);
		}
		flush(f, out, true);
		fclose(f);
	}
}

// -fastx_uniques (derepfull.cpp:214-236): the grouping runs on the device, the files are written on the host
uint64_t FastxUniques(const std::string &InputFileName, const UniquesOpts &Opts)
{
	SeqDB Input;
	Input.FromFasta(InputFileName);
	std::vector<unsigned> UniqOf, First, USize;
	if (usb_device_count() <= 0)
		Die("No CUDA device available: this build has no CPU path");
	DerepFullDevice(Input, UniqOf, First, USize);
	WriteUniques(Input, UniqOf, (unsigned)First.size(), Opts);
	return First.size();
}

uint64_t ClusterFast(const std::string &ReadsFileName, const ClusterOpts &Opts)
{
	if (usb_device_count() <= 0)
		Die("No CUDA device available: this build has no CPU search path");
	const bool timing = getenv("USB_TIMING") != nullptr;
	auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	const double t_start = now();
	// the device context, the (empty) index and the searcher come up while the host reads the input
	usb_index *Index = nullptr;
	usb_searcher *Srch = nullptr;
	std::thread device_up([&]() {
		uint64_t zero_off[1] = {0};
		uint8_t none = 0;
		CheckUsb2(usb_index_create(0, &Opts.P, &none, zero_off, 0, &Index), "usb_index_create");
		CheckUsb2(usb_searcher_create(Index, &Opts.P, &Srch), "usb_searcher_create");
	});
	SeqDB Input;
	Input.FromFasta(ReadsFileName);
	const unsigned SeqCount = Input.GetSeqCount();
	if (SeqCount == 0)
		Die("No sequences in input file");
	const double t_parsed = now();

	std::vector<unsigned> UniqOf, First, USize;
	unsigned UsedCount = SeqCount; // cluster_smallmem -sortedby size -minsize: the input ends at the first smaller size
	if (Opts.smallmem) {
		// every sequence is searched, in input order; the order is only checked (clustersmallmem.cpp:88-127)
		const std::string so = Opts.sortedby.empty() ? "length" : Opts.sortedby;
		if (so != "length" && so != "size" && so != "other")
			Die("Invalid -sortedby");
		unsigned PrevL = UINT_MAX, PrevSize = UINT_MAX;
		for (unsigned i = 0; i < SeqCount; ++i) {
			if (so == "length") {
				const unsigned L = Input.GetSeqLength(i);
				if (L > PrevL)
					Die("Not sorted by length, see -sortedby option in manual");
				PrevL = L;
			} else if (so == "size") {
				const unsigned Size = GetSizeFromLabel(Input.GetLabel(i), UINT_MAX);
				if (Opts.minsize_filled && Size < Opts.minsize) {
					UsedCount = i;
					break;
				}
				if (Size > PrevSize)
					Die("Not sorted by size; prev %u >%s", PrevSize, Input.GetLabel(i));
				PrevSize = Size;
			}
		}
		UniqOf.resize(UsedCount);
		First.resize(UsedCount);
		USize.assign(UsedCount, 1);
		for (unsigned i = 0; i < UsedCount; ++i)
			UniqOf[i] = First[i] = i;
	} else
		DerepFullDevice(Input, UniqOf, First, USize);
	const double t_derep = now();
	const unsigned UniqueCount = (unsigned)First.size();
	// members of each unique in input order (CSR)
	std::vector<unsigned> MemberOff(UniqueCount + 1, 0), Members(UsedCount);
	for (unsigned i = 0; i < UsedCount; ++i)
		++MemberOff[UniqOf[i] + 1];
	for (unsigned u = 0; u < UniqueCount; ++u)
		MemberOff[u + 1] += MemberOff[u];
	{
		std::vector<unsigned> cur(MemberOff.begin(), MemberOff.end() - 1);
		for (unsigned i = 0; i < UsedCount; ++i)
			Members[cur[UniqOf[i]]++] = i;
	}
	// sizes: DerepResult::GetSumSizeIn (derepresult.cpp:211-225, size annotations count whether or not
	// -sizein is given: -sort size) and ClusterSink::GetSize (clustersink.cpp:118-143: annotations
	// only with -sizein, and then every label must have one)
	std::vector<unsigned> SumSizeIn(UniqueCount, 0), QSize(UniqueCount, 0);
	for (unsigned u = 0; u < UniqueCount; ++u)
		for (unsigned m = MemberOff[u]; m < MemberOff[u + 1]; ++m) {
			const char *Label = Input.GetLabel(Members[m]);
			SumSizeIn[u] += GetSizeFromLabel(Label, 1);
			QSize[u] += Opts.sizein ? GetSizeFromLabel(Label, UINT_MAX) : 1u;
		}
	// ---- GetSeqOrder
	std::vector<unsigned> Order(UniqueCount);
	for (unsigned u = 0; u < UniqueCount; ++u)
		Order[u] = u;
	if (Opts.smallmem) {
		if (!Opts.sort.empty())
			Die("-sort is an option of -cluster_fast; -cluster_smallmem takes the input order (-sortedby)");
	} else if (Opts.sort == "length" || Opts.sort == "size") {
		std::vector<unsigned> v(UniqueCount);
		for (unsigned u = 0; u < UniqueCount; ++u)
			v[u] = Opts.sort == "length" ? Input.GetSeqLength(First[u]) : SumSizeIn[u];
		QuickSortOrderDesc(v, Order);
	} else if (!Opts.sort.empty() && Opts.sort != "other" && Opts.sort != "user")
		Die("Invalid sort name %s", Opts.sort.c_str());
	if (Opts.sort == "other")
		Die("-cluster_fast does not support -sort other, use -cluster_smallmem");

	// uniques in cluster order, flattened
	std::vector<uint64_t> Off(UniqueCount + 1, 0);
	for (unsigned k = 0; k < UniqueCount; ++k)
		Off[k + 1] = Off[k] + Input.GetSeqLength(First[Order[k]]);
	std::vector<uint8_t> Letters(Off[UniqueCount] + 16);
	{
		const unsigned T = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
		std::vector<std::thread> th;
		for (unsigned j = 0; j < T; ++j)
			th.emplace_back([&, j]() {
				const unsigned a = (unsigned)((uint64_t)UniqueCount * j / T), b = (unsigned)((uint64_t)UniqueCount * (j + 1) / T);
				for (unsigned k = a; k < b; ++k) {
					const unsigned r = First[Order[k]];
					memcpy(Letters.data() + Off[k], Input.GetSeq(r), Input.GetSeqLength(r));
				}
			});
		for (auto &t : th)
			t.join();
	}
	device_up.join();
	// every unique may become a centroid: reserve once instead of growing through the rounds
	CheckUsb2(usb_index_reserve(Index, UniqueCount, Off[UniqueCount]), "usb_index_reserve");

	const double t_ready = now();
	double t_rounds = 0;
	FILE *fUC = nullptr;
	if (!Opts.uc.empty() && !(fUC = fopen(Opts.uc.c_str(), "wb")))
		Die("Cannot create %s", Opts.uc.c_str());
	std::string buf, cp;
	// OutputSink of the cluster searcher: every unique goes through it in search order, with its hit or without one
	std::unique_ptr<OutputSink> Sink;
	{
		const OutputOpts &O = Opts.Out;
		if (!O.userout.empty() || !O.blast6out.empty() || !O.alnout.empty() || !O.fastapairs.empty() || !O.qsegout.empty() ||
		    !O.tsegout.empty() || !O.matched.empty() || !O.notmatched.empty()) {
			OutputOpts OO = O;
			OO.uc.clear();
			OO.nucleo = Opts.P.is_nucleo != 0;
			OO.local = false;
			OO.match = (int)Opts.P.match;
			OO.mismatch = (int)Opts.P.mismatch;
			Sink.reset(new OutputSink(OO));
		}
	}
	HitMgr SinkHM;
	std::vector<unsigned> ClusterSizes, CentroidRead;
	std::vector<uint32_t> cidx;
	unsigned pos = 0;
	uint32_t B = 256;
	uint64_t rounds = 0;
	while (pos < UniqueCount) {
		const uint32_t cap = Opts.max_block;
		const uint32_t n = std::min<uint32_t>(std::min(B, cap), UniqueCount - pos);
		cidx.resize(n);
		uint32_t ncom = 0;
		usb_result *R = nullptr;
		const double tr0 = now();
		CheckUsb2(usb_cluster_round(Srch, Letters.data(), Off.data() + pos, n, &ncom, cidx.data(), &R), "usb_cluster_round");
		t_rounds += now() - tr0;
		++rounds;
		const usb_hit *hits = usb_result_hits(R);
		const uint64_t *qoff = usb_result_query_offsets(R);
		uint64_t n_runs = 0;
		const uint32_t *runs = usb_result_runs(R, &n_runs);
		for (uint32_t q = 0; q < ncom; ++q) {
			const unsigned u = Order[pos + q];
			const unsigned r0 = First[u];
			const unsigned L = Input.GetSeqLength(r0);
			const unsigned c = cidx[q];
			if (Sink) {
				Input.GetSI(r0, SinkHM.m_Query);
				SinkHM.m_Hits.clear();
				if (qoff[q + 1] != qoff[q]) {
					AlignResult AR;
					AR.m_Hit = hits[qoff[q]];
					AR.m_Hit.target = c; // the centroid database is numbered by cluster (userfield clusternr)
					AR.m_Runs = runs + AR.m_Hit.run_off;
					AR.m_Query = SinkHM.m_Query;
					Input.GetSI(CentroidRead[c], AR.m_Target);
					AR.m_Target.m_Index = c;
					AR.m_Nucleo = Opts.P.is_nucleo != 0;
					SinkHM.m_Hits.push_back(AR);
				}
				Sink->OnQueryDone(SinkHM.m_Query, SinkHM);
			}
			if (qoff[q + 1] == qoff[q]) {
				// ClusterSink::OnQueryDone, no hit: new centroid (clustersink.cpp:318-329)
				if (c != ClusterSizes.size())
					Die("internal: centroid index %u != cluster count %zu", c, ClusterSizes.size());
				ClusterSizes.push_back(QSize[u]);
				CentroidRead.push_back(r0);
				if (fUC) {
					appendf2(buf, "S\t%u\t%u\t*\t.\t*\t*\t*\t", c, L);
					buf += Input.GetLabel(r0);
					buf += "\t*\n";
					for (unsigned m = MemberOff[u] + 1; m < MemberOff[u + 1]; ++m) {
						appendf2(buf, "H\t%u\t%u\t100.0\t.\t0\t%u\t=\t", c, L, L);
						buf += Input.GetLabel(Members[m]);
						buf += '\t';
						buf += Input.GetLabel(r0);
						buf += '\n';
					}
				}
			} else {
				ClusterSizes[c] += QSize[u];
				if (fUC) {
					AlignResult AR;
					AR.m_Hit = hits[qoff[q]];
					AR.m_Runs = runs + AR.m_Hit.run_off;
					AR.GetCompressedPath(cp);
					const char *tlabel = Input.GetLabel(CentroidRead[c]);
					for (unsigned m = MemberOff[u]; m < MemberOff[u + 1]; ++m) {
						appendf2(buf, "H\t%u\t%u\t%.1f\t%c\t%u\t%u\t", c, L, AR.GetPctId(), AR.GetQueryStrand(), 0u, 0u);
						buf += cp;
						buf += '\t';
						buf += Input.GetLabel(Members[m]);
						buf += '\t';
						buf += tlabel;
						buf += '\n';
					}
				}
			}
			if (fUC && buf.size() > (1u << 20)) {
				fwrite(buf.data(), 1, buf.size(), fUC);
				buf.clear();
			}
		}
		usb_result_free(R);
		pos += ncom;
		// block size: grow while whole blocks commit, shrink towards the committed prefix otherwise
		if (ncom == n)
			B = std::min<uint32_t>(B * 2, Opts.max_block);
		else
			B = std::max<uint32_t>(64, std::min<uint32_t>(B, 2 * ncom));
	}
	if (Sink)
		Sink->OnAllDone();
	const double t_loop = now();
	const unsigned ClusterCount = (unsigned)ClusterSizes.size();
	if (fUC) {
		for (unsigned c = 0; c < ClusterCount; ++c) {
			appendf2(buf, "C\t%u\t%u\t*\t*\t*\t*\t*\t", c, ClusterSizes[c]);
			buf += Input.GetLabel(CentroidRead[c]);
			buf += "\t*\n";
		}
		fwrite(buf.data(), 1, buf.size(), fUC);
		fclose(fUC);
	}
	if (!Opts.centroids.empty()) {
		FILE *f = fopen(Opts.centroids.c_str(), "wb");
		if (!f)
			Die("Cannot create %s", Opts.centroids.c_str());
		std::vector<unsigned> COrder;
		QuickSortOrderDesc(ClusterSizes, COrder);
		// clustersink.cpp:245-246: sizes descend, the output stops at the first cluster below -minsize
		unsigned NOut = ClusterCount;
		for (unsigned k = 0; k < ClusterCount; ++k)
			if (ClusterSizes[COrder[k]] < Opts.minsize) {
				NOut = k;
				break;
			}
		// formatted by several threads (one string per slice), written in order
		const unsigned T = std::max(1u, std::min(16u, std::min(std::thread::hardware_concurrency(), NOut / 4096 + 1)));
		std::vector<std::string> parts(T);
		{
			std::vector<std::thread> th;
			for (unsigned j = 0; j < T; ++j)
				th.emplace_back([&, j]() {
					const unsigned k0 = (unsigned)((uint64_t)NOut * j / T), k1 = (unsigned)((uint64_t)NOut * (j + 1) / T);
					std::string &out = parts[j];
					for (unsigned k = k0; k < k1; ++k) {
						const unsigned r = CentroidRead[COrder[k]];
						// ClusterSink::MakeCentroidLabel (clustersink.cpp:217-241); the relabel counter follows the output order
						std::string Label = Input.GetLabel(r);
						if (Opts.sizein || Opts.sizeout)
							StripAnnot(Label, "size=");
						if (!Opts.relabel.empty())
							Label = Opts.relabel + std::to_string(k + 1);
						if (Opts.sizeout)
							AppendSize(Label, ClusterSizes[COrder[k]]);
						out += '>';
						out += Label;
						out += '\n';
						const uint8_t *sq = Input.GetSeq(r);
						const unsigned L = Input.GetSeqLength(r);
						for (unsigned i = 0; i < L; i += 80) {
							out.append((const char *)sq + i, std::min(80u, L - i));
							out += '\n';
						}
					}
				});
			for (auto &t : th)
				t.join();
		}
		for (const std::string &out : parts)
			fwrite(out.data(), 1, out.size(), f);
		std::string out;
		fwrite(out.data(), 1, out.size(), f);
		fclose(f);
	}
	if (!Opts.quiet)
		fprintf(stderr, "%u seqs, %u uniques, %u clusters, %llu rounds\n", SeqCount, UniqueCount, ClusterCount,
		  (unsigned long long)rounds);
	if (timing)
		fprintf(stderr, "timing: parse %.2fs, derep %.2fs, order+flatten+index %.2fs, rounds %.2fs (+ .uc lines %.2fs), C records + centroids %.2fs\n",
		  t_parsed - t_start, t_derep - t_parsed, t_ready - t_derep, t_rounds, t_loop - t_ready - t_rounds, now() - t_loop);
	usb_searcher_free(Srch);
	usb_index_free(Index);
	return ClusterCount;
}

} // namespace usbhost
