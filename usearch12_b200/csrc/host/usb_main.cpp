// usb_main.cpp -- command line of the GPU hot path, with the reference's option syntax for the
// commands on the path (usearch_main.cpp:19-71, opts.cpp:272-362: "-opt value" or "--opt value").
//
//   usearch12_b200 -usearch_global Q.fa -db DB.fa -id 0.97 -strand plus|both
//        [-maxaccepts n] [-maxrejects n] [-uc f] [-blast6out f] [-userout f] [-userfields a+b+..]
//        [-dbmask fastnucleo|none] [-gpus n] [-threads n] [-quiet]
// Options of the reference that this build does not implement are refused (Die), never ignored.
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>

#include "usb_host.h"

using namespace usbhost;

int main(int argc, char **argv)
{
	std::map<std::string, std::string> opt;
	static const char *flags[] = {"quiet", "output_no_hits", "fulldp", nullptr};
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
	const std::string reads = take("cluster_fast", nullptr);
	if (!reads.empty()) {
		// clusterfast.cpp:135 cmd_cluster_fast
		ClusterOpts C;
		usb_default_params(&C.P, 1);
		const std::string cid = take("id", nullptr);
		if (cid.empty())
			Die("Must specify -id"); // makeclustersearcher.cpp:30-31
		C.P.id = (float)atof(cid.c_str());
		const std::string cstrand = take("strand", "plus");
		if (cstrand != "plus")
			Die("-cluster_fast -strand %s is not supported by this build", cstrand.c_str());
		C.P.maxrejects = (uint32_t)atoi(take("maxrejects", "8").c_str());
		C.uc = take("uc", nullptr);
		C.centroids = take("centroids", nullptr);
		C.sort = take("sort", nullptr);
		C.max_block = (uint32_t)atoi(take("batch", "65536").c_str());
		C.quiet = !take("quiet", nullptr).empty();
		take("threads", nullptr); // derep/cluster order follow the reference's -threads 1 behaviour
		if (!opt.empty())
			Die("Option -%s is not supported by this build", opt.begin()->first.c_str());
		ClusterFast(reads, C);
		return 0;
	}
	SearchOpts O;
	usb_default_params(&O.P, 0);
	const std::string query = take("usearch_global", nullptr);
	if (query.empty())
		Die("No command: this build implements -usearch_global and -cluster_fast");
	const std::string db = take("db", nullptr);
	const std::string id = take("id", nullptr);
	if (id.empty())
		Die("-id option required"); // accepter.cpp: identity threshold is mandatory for usearch_global
	O.P.id = (float)atof(id.c_str());
	const std::string strand = take("strand", nullptr);
	if (strand != "plus" && strand != "both")
		Die("Must specify -strand plus or both with nt db"); // search.cpp:23-34
	O.P.strand_both = strand == "both";
	O.P.maxaccepts = (uint32_t)atoi(take("maxaccepts", "1").c_str());
	O.P.maxrejects = (uint32_t)atoi(take("maxrejects", "32").c_str());
	O.P.band = (uint32_t)atoi(take("band", "16").c_str());   // alnheuristics.cpp:33
	O.P.fulldp = !take("fulldp", nullptr).empty();            // alnheuristics.cpp:64-76
	const std::string dbmask = take("dbmask", "fastnucleo");
	if (dbmask != "fastnucleo" && dbmask != "none")
		Die("-dbmask %s not supported (fastnucleo|none)", dbmask.c_str());
	O.P.dbmask = dbmask == "fastnucleo";
	O.Out.uc = take("uc", nullptr);
	O.Out.blast6out = take("blast6out", nullptr);
	O.Out.userout = take("userout", nullptr);
	O.Out.userfields = take("userfields", nullptr);
	O.Out.output_no_hits = !take("output_no_hits", nullptr).empty();
	O.gpus = atoi(take("gpus", "1").c_str());
	O.batch = (uint32_t)atoi(take("batch", "262144").c_str());
	O.quiet = !take("quiet", nullptr).empty();
	take("threads", nullptr); // host threads are not on the search path here
	if (!opt.empty())
		Die("Option -%s is not supported by this build", opt.begin()->first.c_str());
	const uint64_t n = Search(query, db, O);
	if (!O.quiet)
		fprintf(stderr, "%llu queries matched\n", (unsigned long long)n);
	return 0;
}
