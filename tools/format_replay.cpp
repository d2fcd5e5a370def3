// format_replay.cpp -- TEST TOOL (not part of the product): drives the host-side sinks of
// usearch12_b200/csrc/host (OutputSink, DBHitSink) with hits read from a table instead of hits
// computed on the device, so that every output format can be compared byte for byte with files the
// unmodified reference binary wrote, on a machine without a GPU (tests/test_formats_cpu.py).
//
// The table is the reference's own -userout with the fields
//   query+target+qstrand+ids+mism+gaps+opens+qlot+qhit+tlot+thit+alnlen+ql+tl+raw+aln
// (tools/make_golden_formats.py), i.e. exactly the numbers a usb_hit carries.  Nothing is aligned or
// recomputed here; the formats under test get the same inputs as in a real run.
//
//   format_replay -query Q.fa -db DB.udb|DB.fa -hits HITS.tsv [-local 1 -evalue E] [-amino 1]
//                 [-uc f] [-blast6out f] [-userout f -userfields a+b] [-alnout f] [-fastapairs f] [-qsegout f]
//                 [-tsegout f] [-matched f] [-notmatched f] [-matchedfq f] [-notmatchedfq f] [-dbmatched f] [-dbnotmatched f] [-sizein] [-sizeout]
#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "../usearch12_b200/csrc/host/usb_host.h"

using namespace usbhost;

static std::vector<std::string> Split(const std::string &s, char sep)
{
	std::vector<std::string> f;
	size_t pos = 0;
	for (;;) {
		size_t e = s.find(sep, pos);
		if (e == std::string::npos) {
			f.push_back(s.substr(pos));
			return f;
		}
		f.push_back(s.substr(pos, e - pos));
		pos = e + 1;
	}
}

int main(int argc, char **argv)
{
	std::map<std::string, std::string> opt;
	for (int i = 1; i < argc; ++i) {
		const char *a = argv[i];
		while (*a == '-')
			++a;
		if (!strcmp(a, "sizein") || !strcmp(a, "sizeout") || !strcmp(a, "output_no_hits") || !strcmp(a, "uc_hitsonly") || !strcmp(a, "top_hit_only") || !strcmp(a, "top_hits_only"))
			opt[a] = "1";
		else if (i + 1 < argc)
			opt[a] = argv[++i];
		else
			Die("Missing value for -%s", a);
	}
	auto get = [&](const char *k) { return opt.count(k) ? opt[k] : std::string(); };
	if (!get("uniques").empty()) {
		// -uniques IN.fa: the writers of -fastx_uniques (WriteUniques) behind a grouping made here with a map
		// (test only; the product groups on the device: usb_derep_full)
		SeqDB In;
		In.FromFasta(get("uniques"));
		std::unordered_map<std::string, unsigned> seen;
		std::vector<unsigned> UniqOf(In.GetSeqCount());
		for (uint32_t i = 0; i < In.GetSeqCount(); ++i) {
			std::string s((const char *)In.GetSeq(i), In.GetSeqLength(i));
			for (char &c : s)
				c = (char)toupper((unsigned char)c);
			UniqOf[i] = seen.emplace(s, (unsigned)seen.size()).first->second;
		}
		UniquesOpts U;
		U.fastaout = get("fastaout");
		U.relabel = get("relabel");
		U.sizein = !get("sizein").empty();
		U.sizeout = !get("sizeout").empty();
		U.topn = (unsigned)atoi(get("topn").c_str());
		U.minuniquesize = (unsigned)atoi(get("minuniquesize").c_str());
		WriteUniques(In, UniqOf, (unsigned)seen.size(), U);
		return 0;
	}
	const bool local = get("local") == "1", amino = get("amino") == "1";
	usb_params P;
	usb_default_params(&P, 0);
	if (local)
		usb_set_local(&P, amino ? 0 : 1, (float)atof(get("evalue").c_str()));
	else if (amino)
		usb_set_amino(&P);

	SeqDB Q, DB;
	Q.FromFasta(get("query"));
	Q.DropSmallerThan((unsigned)atoi(get("minsize").c_str())); // as Search() does before the first batch
	if (IsUDBFile(get("db"))) {
		bool nucleo = true;
		uint32_t wl = 0;
		DB.FromUDB(get("db"), nucleo, wl); // letters masked as the reference's database holds them
	} else
		DB.FromFasta(get("db"));
	// query labels need not be unique: the table is in query order (the reference ran with one thread), so a hit
	// belongs to the first query of that label at or behind the query of the hit before it
	std::unordered_map<std::string, std::vector<uint32_t>> qidx;
	std::unordered_map<std::string, uint32_t> tidx;
	uint32_t cursor = 0;
	for (uint32_t i = 0; i < Q.GetSeqCount(); ++i)
		qidx[Q.GetLabel(i)].push_back(i);
	for (uint32_t i = 0; i < DB.GetSeqCount(); ++i)
		tidx.emplace(DB.GetLabel(i), i);

	// hits of every query, in file order = the order HitMgr handed them to the reference's sinks
	struct Row {
		usb_hit h;
		std::vector<uint32_t> runs;
	};
	std::vector<std::vector<Row>> rows(Q.GetSeqCount());
	std::ifstream in(get("hits"));
	if (!in)
		Die("Cannot open %s", get("hits").c_str());
	std::string line;
	while (std::getline(in, line)) {
		if (line.empty())
			continue;
		const std::vector<std::string> f = Split(line, '\t');
		if (f.size() != 16)
			Die("hit table: %u fields, 16 expected", (unsigned)f.size());
		auto qi = qidx.find(f[0]);
		auto ti = tidx.find(f[1]);
		if (qi == qidx.end() || ti == tidx.end())
			Die("hit table: unknown label %s / %s", f[0].c_str(), f[1].c_str());
		Row r;
		memset(&r.h, 0, sizeof r.h);
		usb_hit &h = r.h;
		{
			const std::vector<uint32_t> &ids = qi->second;
			auto it = std::lower_bound(ids.begin(), ids.end(), cursor);
			if (it == ids.end())
				Die("hit table: hits of %s are not in query order", f[0].c_str());
			h.query = cursor = *it;
		}
		h.target = ti->second;
		h.strand = f[2] == "-";
		h.ids = (uint32_t)atoi(f[3].c_str());
		h.mism = (uint32_t)atoi(f[4].c_str());
		h.intgaps = (uint32_t)atoi(f[5].c_str());
		h.opens = (uint32_t)atoi(f[6].c_str());
		h.first_mq = (uint32_t)atoi(f[7].c_str());
		h.last_mq = (uint32_t)atoi(f[8].c_str());
		h.first_mt = (uint32_t)atoi(f[9].c_str());
		h.last_mt = (uint32_t)atoi(f[10].c_str());
		h.alnlen = (uint32_t)atoi(f[11].c_str());
		h.ql = (uint32_t)atoi(f[12].c_str());
		h.tl = (uint32_t)atoi(f[13].c_str());
		h.raw = (int32_t)atoi(f[14].c_str());
		const std::string &path = f[15];
		h.first_mcol = (uint32_t)path.find('M');
		for (size_t i = 0; i < path.size();) {
			size_t j = i;
			while (j < path.size() && path[j] == path[i])
				++j;
			const uint32_t op = path[i] == 'M' ? 0u : path[i] == 'D' ? 1u : 2u;
			r.runs.push_back((uint32_t)((j - i) << 2) | op);
			i = j;
		}
		h.run_off = 0;
		h.run_cnt = (uint32_t)r.runs.size();
		rows[h.query].push_back(std::move(r));
	}

	OutputOpts O;
	O.uc = get("uc");
	O.blast6out = get("blast6out");
	O.userout = get("userout");
	O.userfields = get("userfields");
	O.alnout = get("alnout");
	O.fastapairs = get("fastapairs");
	O.qsegout = get("qsegout");
	O.tsegout = get("tsegout");
	O.matched = get("matched");
	O.notmatched = get("notmatched");
	O.trimout = get("trimout");
	O.matchedfq = get("matchedfq");
	O.notmatchedfq = get("notmatchedfq");
	if (!get("rowlen").empty())
		O.rowlen = (unsigned)atoi(get("rowlen").c_str());
	if (!get("flank").empty())
		O.flank = (unsigned)atoi(get("flank").c_str());
	O.output_no_hits = !get("output_no_hits").empty();
	O.uc_hitsonly = !get("uc_hitsonly").empty();
	O.cmdline = "format_replay ";
	O.nucleo = !amino;
	O.local = local;
	O.match = (int)P.match;
	O.mismatch = (int)P.mismatch;
	HitSelection Sel;
	Sel.maxhits = (unsigned)atoi(get("maxhits").c_str());
	Sel.top_hit_only = !get("top_hit_only").empty();
	Sel.top_hits_only = !get("top_hits_only").empty();
	OutputSink Sink(O);
	DBHitSink DbSink(DB, get("dbmatched"), get("dbnotmatched"), !get("sizein").empty(), !get("sizeout").empty(), get("dbcutout"),
	  !get("otutabout").empty() || !get("mapout").empty() || !get("biomout").empty());
	// -otutab: the OTU table sink behind the same hit lists (otutabsink.cpp:25-76)
	std::unique_ptr<OtuTabSink> OtuSink;
	if (!get("otutabout").empty() || !get("mapout").empty() || !get("biomout").empty())
		OtuSink.reset(new OtuTabSink(get("otutabout"), get("mapout"), get("sample_delim"), true, get("biomout")));

	// batches of 5000 queries, so that both the serial and the threaded formatting paths run
	const uint32_t NQ = Q.GetSeqCount(), B = (uint32_t)std::max(1, atoi(get("batch").empty() ? "5000" : get("batch").c_str()));
	for (uint32_t first = 0; first < NQ; first += B) {
		const uint32_t count = std::min(B, NQ - first);
		std::vector<HitMgr> batch(count);
		for (uint32_t q = 0; q < count; ++q) {
			HitMgr &HM = batch[q];
			Q.GetSI(first + q, HM.m_Query);
			for (const Row &r : rows[first + q]) {
				AlignResult AR;
				AR.m_Hit = r.h;
				AR.m_Query = HM.m_Query;
				AR.m_Query.m_RevComp = r.h.strand != 0;
				DB.GetSI(r.h.target, AR.m_Target);
				AR.m_Runs = r.runs.data();
				AR.m_Nucleo = !amino;
				AR.m_Local = local;
				if (local && usb_params_evalue(&P, r.h.raw, r.h.ql, &AR.m_Evalue, &AR.m_BitScore) != 0)
					Die("usb_params_evalue: %s", usb_last_error());
				HM.m_Hits.push_back(AR);
			}
			SelectHits(HM.m_Hits, Sel); // -maxhits / -top_hit_only / -top_hits_only, as BuildHitMgrs does
		}
		Sink.OnBatchDone(batch);
		DbSink.OnBatchDone(batch);
		if (OtuSink)
			OtuSink->OnBatchDone(batch);
	}
	Sink.OnAllDone();
	DbSink.OnAllDone();
	if (OtuSink)
		OtuSink->OnAllDone();
	return 0;
}
