// usb_main.cpp -- command line of the GPU hot path, with the reference's option syntax for the
// commands on the path (usearch_main.cpp:19-71, opts.cpp:272-362: "-opt value" or "--opt value").
//
//   usearch12_b200 -usearch_global Q.fa -db DB.fa -id 0.97 -strand plus|both
//        [-maxaccepts n] [-maxrejects n] (0 = no limit) [-uc f] [-blast6out f] [-userout f] [-userfields a+b+..]
//        [-alnout f] [-fastapairs f] [-qsegout f] [-tsegout f] [-matched f] [-notmatched f] [-matchedfq f]
//        [-notmatchedfq f] [-trimout f] [-dbmatched f] [-dbnotmatched f] [-dbcutout f] [-sizein] [-sizeout]
//        [-rowlen n] [-flank n] [-uc_hitsonly] [-output_no_hits] [-minsize n]
//        [-match x] [-mismatch x] [-minhsp n] [-xdrop_nw x] [-hspw n] [-band n] [-bump n] [-fulldp]
//        [-dbmask fastnucleo|none] [-gpus n] [-threads n] [-quiet]      (queries: FASTA or FASTQ)
//   usearch12_b200 -otutab READS.fa -otus|-zotus|-db OTUS.fa -otutabout TABLE.txt [-mapout MAP.txt]
//        [-sample_delim s] (searchcmd.cpp:21-40: -id 0.97 -strand both -maxaccepts 3 -maxrejects 32)
//   usearch12_b200 -fastx_uniques IN.fa -fastaout OUT.fa [-sizeout] [-relabel prefix] [-minuniquesize n]
//   usearch12_b200 -makeudb_usearch DB.fa -output DB.udb [-dbmask fastnucleo|fastamino|none]
//        (the file is byte for byte the reference's; -db accepts FASTA or .udb, from either program)
//   usearch12_b200 -usearch_local Q.fa -db DB.fa -id 0.5 -evalue 1e-5 [-strand plus|both for nt DBs]
//        [-maxaccepts n] [-maxrejects n] [-xdrop_u x] [-xdrop_g x] [-lopen x] [-lext x] [-ka_dbsize x] ...
// Options of the reference that this build does not implement are refused (Die), never ignored.
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>

#include "usb_host.h"

using namespace usbhost;

// StrToUint / StrToFloat (myutils.cpp:1148-1155,1217-1231): a malformed number stops the program
static unsigned ToUint(const std::string &s)
{
	if (s.empty() || s.find_first_not_of("0123456789") != std::string::npos)
		Die("Invalid integer '%s'", s.c_str());
	return (unsigned)strtoul(s.c_str(), nullptr, 10);
}

static double ToFloat(const std::string &s)
{
	char *end = nullptr;
	const double d = strtod(s.c_str(), &end);
	if (s.empty() || !end || *end != 0)
		Die("Invalid floating-point number '%s'", s.c_str());
	return d;
}

int main(int argc, char **argv)
{
	std::map<std::string, std::string> opt;
	static const char *flags[] = {"quiet", "output_no_hits", "fulldp", "sizeout", "sizein", "self", "notself", "selfid", "top_hit_only",
	                              "top_hits_only", "uc_hitsonly", nullptr};
	for (int i = 1; i < argc; ++i) {
		const char *a = argv[i];
		if (a[0] != '-')
			Die("Unexpected argument '%s'", a);
		while (*a == '-')
			++a;
		bool is_flag = false;
		for (const char **f = flags; *f; ++f)
			is_flag |= strcmp(*f, a) == 0;
		if (is_flag)
			opt[a] = "1";
		else {
			if (i + 1 >= argc)
				Die("Missing value for option -%s", a);
			opt[a] = argv[++i];
		}
	}
	auto take = [&](const char *name, const char *dflt) -> std::string {
		auto it = opt.find(name);
		if (it == opt.end())
			return dflt ? dflt : "";
		std::string v = it->second;
		opt.erase(it);
		return v;
	};
	std::string reads = take("cluster_fast", nullptr);
	const std::string reads_sm = take("cluster_smallmem", nullptr);
	if (!reads.empty() || !reads_sm.empty()) {
		// clusterfast.cpp:135 cmd_cluster_fast, clustersmallmem.cpp:145 cmd_cluster_smallmem
		ClusterOpts C;
		usb_default_params(&C.P, 1);
		if (reads.empty()) {
			reads = reads_sm;
			C.smallmem = true;
			C.sortedby = take("sortedby", nullptr);
			if (opt.count("fastaout"))
				Die("-fastaout not supported, use -centroids"); // clustersmallmem.cpp:55-61
		}
		const std::string cid = take("id", nullptr);
		if (cid.empty())
			Die("Must specify -id"); // makeclustersearcher.cpp:30-31
		C.P.id = (float)ToFloat(cid);
		const std::string cstrand = take("strand", "plus");
		if (cstrand != "plus")
			Die("-cluster_fast -strand %s is not supported by this build", cstrand.c_str());
		C.P.maxrejects = (uint32_t)ToUint(take("maxrejects", C.smallmem ? "32" : "8")); // terminator.cpp:10-31
		C.uc = take("uc", nullptr);
		C.centroids = take("centroids", nullptr);
		C.sort = take("sort", nullptr);
		C.Out.userout = take("userout", nullptr);
		C.Out.userfields = take("userfields", nullptr);
		C.Out.blast6out = take("blast6out", nullptr);
		C.Out.alnout = take("alnout", nullptr);
		C.Out.fastapairs = take("fastapairs", nullptr);
		C.Out.qsegout = take("qsegout", nullptr);
		C.Out.tsegout = take("tsegout", nullptr);
		C.Out.matched = take("matched", nullptr);
		C.Out.notmatched = take("notmatched", nullptr);
		C.Out.rowlen = (unsigned)ToUint(take("rowlen", "80"));
		if (C.Out.rowlen == 0)
			Die("-rowlen must be positive");
		for (int i = 0; i < argc; ++i) {
			C.Out.cmdline += argv[i];
			C.Out.cmdline += ' ';
		}
		C.sizein = !take("sizein", nullptr).empty();
		C.sizeout = !take("sizeout", nullptr).empty();
		C.relabel = take("relabel", nullptr);
		C.minsize_filled = opt.count("minsize") != 0;
		C.minsize = (unsigned)ToUint(take("minsize", "0"));
		C.max_block = (uint32_t)ToUint(take("batch", "65536"));
		C.quiet = !take("quiet", nullptr).empty();
		take("threads", nullptr); // derep/cluster order follow the reference's -threads 1 behaviour
		if (!opt.empty())
			Die("Option -%s is not supported by this build", opt.begin()->first.c_str());
		ClusterFast(reads, C);
		return 0;
	}
	const std::string un = take("fastx_uniques", nullptr);
	if (!un.empty()) {
		// derepfull.cpp:233 cmd_fastx_uniques (host only)
		UniquesOpts U;
		U.fastaout = take("fastaout", nullptr);
		U.relabel = take("relabel", nullptr);
		U.sizeout = !take("sizeout", nullptr).empty();
		U.sizein = !take("sizein", nullptr).empty();
		U.topn = (unsigned)ToUint(take("topn", "0"));
		U.minuniquesize = (unsigned)ToUint(take("minuniquesize", "0"));
		const bool quiet = !take("quiet", nullptr).empty();
		take("threads", nullptr); // the unique order follows the reference's -threads 1 behaviour
		if (opt.count("output"))
			Die("Use -fastaout, not -output"); // derepfull.cpp:216-217
		if (!opt.empty())
			Die("Option -%s is not supported by this build", opt.begin()->first.c_str());
		const uint64_t n = FastxUniques(un, U);
		if (!quiet)
			fprintf(stderr, "%llu uniques\n", (unsigned long long)n);
		return 0;
	}
	const std::string mk = take("makeudb_usearch", nullptr);
	if (!mk.empty()) {
		// makeudb.cpp:62 cmd_makeudb_usearch
		usb_params P;
		usb_default_params(&P, 0);
		const bool nt = GuessIsNucleo(mk);
		if (!nt)
			usb_set_local(&P, 0, 1.0f); // amino acid alphabet: words of 5 over 20 letters
		const std::string mask = take("dbmask", nt ? "fastnucleo" : "fastamino");
		if (mask != (nt ? "fastnucleo" : "fastamino") && mask != "none")
			Die("-dbmask %s not supported (%s|none)", mask.c_str(), nt ? "fastnucleo" : "fastamino");
		P.dbmask = mask != "none";
		const std::string wl = take("wordlength", nullptr); // udbparams.cpp:58-81: index words of another length
		if (!wl.empty()) {
			P.word_length = (uint32_t)ToUint(wl);
			if (P.word_length < 2 || P.word_length > (nt ? 8u : 5u))
				Die("-wordlength %s not supported (%s)", wl.c_str(), nt ? "2..8" : "2..5");
		}
		const std::string out = take("output", nullptr);
		take("quiet", nullptr);
		take("threads", nullptr);
		if (!opt.empty())
			Die("Option -%s is not supported by this build", opt.begin()->first.c_str());
		MakeUDB(mk, out, P);
		return 0;
	}
	SearchOpts O;
	usb_default_params(&O.P, 0);
	std::string query = take("usearch_global", nullptr);
	const std::string lquery = take("usearch_local", nullptr);
	const std::string oquery = take("otutab", nullptr);
	const std::string cquery = take("closed_ref", nullptr);
	std::unique_ptr<ClosedRefSink> crsink;
	if (query.empty() && lquery.empty() && oquery.empty() && !cquery.empty()) {
		// searchcmd.cpp:11-19 cmd_closed_ref: -id 0.97 and -stepwords 0 unless given, Terminator 4 / 16
		// (terminator.cpp:16-20), GlobalAligner, ClosedRefSink behind the OutputSink
		query = cquery;
	}
	if (query.empty() && lquery.empty() && oquery.empty())
		Die("No command: this build implements -usearch_global, -usearch_local, -otutab, -cluster_fast and -makeudb_usearch");
	std::string db = take("db", nullptr);
	std::string id = take("id", nullptr);
	if (!cquery.empty()) {
		if (id.empty())
			id = "0.97";
		O.P.stepwords = 0;
		crsink.reset(new ClosedRefSink(take("tabbedout", nullptr), take("dbotus", nullptr), take("dataotus", nullptr)));
		O.ExtraSinks.push_back(crsink.get());
		O.ClosedRef = crsink.get();
	}
	std::unique_ptr<OtuTabSink> otusink;
	const char *dflt_strand = nullptr, *dflt_ma = "1", *dflt_mr = "32";
	if (!cquery.empty()) { // terminator.cpp:16-20
		dflt_ma = "4";
		dflt_mr = "16";
	}
	if (!oquery.empty()) {
		// searchcmd.cpp:21-40 cmd_otutab: defaults, then the OTU FASTA from -db, -otus or -zotus
		query = oquery;
		if (id.empty())
			id = "0.97";
		dflt_strand = "both";
		dflt_ma = "3";
		const std::string otus = take("otus", nullptr), zotus = take("zotus", nullptr);
		if (db.empty())
			db = !otus.empty() ? otus : zotus;
		if (db.empty())
			Die("Must specify OTU FASTA -db, -otus or -zotus");
		const std::string tab = take("otutabout", nullptr), mapout = take("mapout", nullptr);
		const std::string biom = take("biomout", nullptr);
		otusink.reset(new OtuTabSink(tab, mapout, take("sample_delim", nullptr), opt.count("quiet") != 0, biom));
		O.ExtraSinks.push_back(otusink.get());
		O.otutab = true;
	}
	if (id.empty())
		Die("--id not set"); // udbusortedsearcher.cpp:99-100: mandatory for both commands
	O.P.id = (float)ToFloat(id);
	bool nucleo = true;
	if (!lquery.empty()) {
		// searchcmd.cpp:42 cmd_usearch_local; the DB alphabet is guessed from its letters like
		// SeqDB::GetIsNucleo (seqdb.cpp:181-199): more than 90 % of the first letters are ACGTUN
		query = lquery;
		const std::string ev = take("evalue", nullptr);
		if (ev.empty())
			Die("Must set -evalue"); // accepter.cpp / search.cpp: mandatory for local searches
		nucleo = IsUDBFile(db) ? UDBIsNucleo(db) : GuessIsNucleo(db);
		usb_set_local(&O.P, nucleo ? 1 : 0, (float)ToFloat(ev));
		O.P.xdrop_u = (float)ToFloat(take("xdrop_u", "16"));
		O.P.xdrop_g = (float)ToFloat(take("xdrop_g", "32"));
		const std::string lo = take("lopen", nullptr), le = take("lext", nullptr);
		if (lo.empty() != le.empty())
			Die("Must set both --lopen and --lext"); // alnparams.cpp:362-366
		if (!lo.empty()) {
			if (ToFloat(lo) < 0.0 || ToFloat(le) < 0.0)
				Die("Invalid --lopen/--lext, gap penalties must be >= 0");
			O.P.lopen = -(float)ToFloat(lo);
			O.P.lext = -(float)ToFloat(le);
		}
		O.P.ka_dbsize = (float)ToFloat(take("ka_dbsize", "1e9"));
		const std::string hw = take("hspw", nullptr);
		if (!hw.empty())
			O.P.hspw = (uint32_t)ToUint(hw);
	} else if (oquery.empty()) {
		// -usearch_global takes either alphabet (makedbsearcher.cpp:132-140); amino acid databases
		// switch to BLOSUM62, gap open -17 and HSP words of 3 letters
		// (a missing database is reported by Search() after the option checks below)
		FILE *probe = fopen(db.c_str(), "rb");
		if (probe) {
			fclose(probe);
			nucleo = IsUDBFile(db) ? UDBIsNucleo(db) : GuessIsNucleo(db);
		}
		if (!nucleo)
			usb_set_amino(&O.P);
	}
	const std::string strand = take("strand", dflt_strand);
	if (nucleo && strand != "plus" && strand != "both")
		Die("Must specify -strand plus or both with nt db"); // search.cpp:23-34
	O.P.strand_both = nucleo && strand == "both";
	O.P.maxaccepts = (uint32_t)ToUint(take("maxaccepts", dflt_ma));
	O.P.maxrejects = (uint32_t)ToUint(take("maxrejects", dflt_mr));
	if (lquery.empty() && nucleo) {
		// nucleotide substitution scores (alnparams.cpp:330-334 SetNucSubstMx) and the HSP heuristics of the
		// global aligner (alnheuristics.cpp:26-44); a local search keeps its defaults here
		O.P.match = (float)ToFloat(take("match", "1"));
		O.P.mismatch = (float)ToFloat(take("mismatch", "-2"));
		O.P.minhsp = (uint32_t)ToUint(take("minhsp", "16"));
		O.P.xdrop_nw = (float)ToFloat(take("xdrop_nw", "8"));
		const std::string hw = take("hspw", nullptr); // alnheuristics.cpp:60-61
		if (!hw.empty())
			O.P.hspw = (uint32_t)ToUint(hw);
	}
	O.P.bump = (uint32_t)ToUint(take("bump", "50")); // udbusortedsearcher.cpp:269-282
	O.P.big = (uint32_t)ToUint(take("big", "100000")); // udbusortedsearcher.cpp:39-58: UDBSearchBig above this many targets
	O.P.band = (uint32_t)ToUint(take("band", "16"));   // alnheuristics.cpp:33
	O.P.fulldp = !take("fulldp", nullptr).empty();            // alnheuristics.cpp:64-76
	const std::string dbmask = take("dbmask", nucleo ? "fastnucleo" : "fastamino");
	if (dbmask != (nucleo ? "fastnucleo" : "fastamino") && dbmask != "none")
		Die("-dbmask %s not supported (%s|none)", dbmask.c_str(), nucleo ? "fastnucleo" : "fastamino");
	O.P.dbmask = dbmask != "none";
	O.Out.uc = take("uc", nullptr);
	O.Out.blast6out = take("blast6out", nullptr);
	O.Out.userout = take("userout", nullptr);
	O.Out.userfields = take("userfields", nullptr);
	O.Out.output_no_hits = !take("output_no_hits", nullptr).empty();
	O.Out.uc_hitsonly = !take("uc_hitsonly", nullptr).empty();
	O.minsize = (unsigned)ToUint(take("minsize", "0"));
	// the other files of OutputSink::OpenOutputFiles (outputsink.cpp:135-195) and of DBHitSink (dbhitsink.cpp:42-50)
	O.Out.alnout = take("alnout", nullptr);
	O.Out.fastapairs = take("fastapairs", nullptr);
	O.Out.qsegout = take("qsegout", nullptr);
	O.Out.tsegout = take("tsegout", nullptr);
	O.Out.matched = take("matched", nullptr);
	O.Out.notmatched = take("notmatched", nullptr);
	O.Out.trimout = take("trimout", nullptr);
	O.Out.matchedfq = take("matchedfq", nullptr);
	O.Out.notmatchedfq = take("notmatchedfq", nullptr);
	O.Out.rowlen = (unsigned)ToUint(take("rowlen", "80"));
	if (O.Out.rowlen == 0)
		Die("-rowlen must be positive");
	O.Out.flank = (unsigned)ToUint(take("flank", "8"));
	O.dbmatched = take("dbmatched", nullptr);
	O.dbnotmatched = take("dbnotmatched", nullptr);
	O.dbcutout = take("dbcutout", nullptr);
	if (cquery.empty()) { // DBHitSink: -sizein weighs a hit with the query's size=, -sizeout annotates the matched targets
		O.sizein = !take("sizein", nullptr).empty();
		O.sizeout = !take("sizeout", nullptr).empty();
	}
	for (int i = 0; i < argc; ++i) { // PrintCmdLine (myutils.cpp:1667-1674): every argument followed by a blank
		O.Out.cmdline += argv[i];
		O.Out.cmdline += ' ';
	}
	// Accepter / Terminator / HitMgr options (accepter.cpp:41-94,145-197; terminator.cpp:66-86;
	// hitmgr.cpp:367-398)
	{
		auto flag = [&](const char *name, uint32_t bit) {
			if (!take(name, nullptr).empty())
				O.P.accept_flags |= bit;
		};
		auto flt = [&](const char *name, uint32_t bit, float &dst) {
			const std::string v = take(name, nullptr);
			if (!v.empty()) {
				dst = (float)ToFloat(v);
				O.P.accept_flags |= bit;
			}
		};
		auto uns = [&](const char *name, uint32_t bit, uint32_t &dst) {
			const std::string v = take(name, nullptr);
			if (!v.empty()) {
				dst = (uint32_t)ToUint(v);
				O.P.accept_flags |= bit;
			}
		};
		flag("self", USB_ACC_SELF);
		flag("notself", USB_ACC_NOTSELF);
		flag("selfid", USB_ACC_SELFID);
		flt("maxid", USB_ACC_MAXID, O.P.maxid);
		uns("mincols", USB_ACC_MINCOLS, O.P.mincols);
		uns("maxgaps", USB_ACC_MAXGAPS, O.P.maxgaps);
		flt("query_cov", USB_ACC_QUERY_COV, O.P.query_cov);
		flt("max_query_cov", USB_ACC_MAX_QUERY_COV, O.P.max_query_cov);
		flt("target_cov", USB_ACC_TARGET_COV, O.P.target_cov);
		flt("max_target_cov", USB_ACC_MAX_TARGET_COV, O.P.max_target_cov);
		uns("maxdiffs", USB_ACC_MAXDIFFS, O.P.maxdiffs);
		uns("mindiffs", USB_ACC_MINDIFFS, O.P.mindiffs);
		flt("abskew", USB_ACC_ABSKEW, O.P.abskew);
		flt("min_sizeratio", USB_ACC_MIN_SIZERATIO, O.P.min_sizeratio);
		flt("minqt", USB_ACC_MINQT, O.P.minqt);
		flt("maxqt", USB_ACC_MAXQT, O.P.maxqt);
		flt("minsl", USB_ACC_MINSL, O.P.minsl);
		flt("maxsl", USB_ACC_MAXSL, O.P.maxsl);
		flt("termid", USB_ACC_TERMID, O.P.termid);
		flt("termidd", USB_ACC_TERMIDD, O.P.termidd);
		const std::string mh = take("maxhits", nullptr);
		if (!mh.empty())
			O.Sel.maxhits = (unsigned)ToUint(mh);
		O.Sel.top_hit_only = !take("top_hit_only", nullptr).empty();
		O.Sel.top_hits_only = !take("top_hits_only", nullptr).empty();
	}
	O.gpus = ToUint(take("gpus", "1"));
	O.batch = (uint32_t)ToUint(take("batch", "262144"));
	O.quiet = !take("quiet", nullptr).empty();
	take("threads", nullptr); // host threads are not on the search path here
	if (!opt.empty())
		Die("Option -%s is not supported by this build", opt.begin()->first.c_str());
	const uint64_t n = Search(query, db, O);
	if (!O.quiet)
		fprintf(stderr, "%llu queries matched\n", (unsigned long long)n);
	return 0;
}
