// ORACLE — test infrastructure only.  `vs_oracle construct|query`: CPU stand-in for the reference
// CLI (src/variantstore.cc:81-156, src/commands.cc:32-60,113-215), used for manual checks and as the
// timed CPU arm of the bench.
#include "vso.h"
#include <chrono>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <sys/stat.h>
using namespace vso;

static const char* arg(int argc, char** argv, const char* f, const char* d = nullptr) {
	for (int i = 2; i + 1 < argc; i++) if (!strcmp(argv[i], f)) return argv[i + 1];
	return d;
}
static bool flag(int argc, char** argv, const char* f) { for (int i = 2; i < argc; i++) if (!strcmp(argv[i], f)) return true; return false; }

int main(int argc, char** argv) {
	if (argc < 2) { fprintf(stderr, "usage: vs_oracle construct -r ref.fa -v x.vcf -p prefix | query -p prefix -t 4|6|7 -r regions [-s sample] [-a alts] [-b refs] [-o out] [-v]\n"); return 1; }
	try {
		std::string cmd = argv[1];
		if (cmd == "construct") {
			ConstructOpts o;
			if (const char* l = arg(argc, argv, "--cqf-log2")) o.cqf_log2_slots = atoi(l);
			std::string prefix = arg(argc, argv, "-p", "ser");
			mkdir(prefix.c_str(), 0755);
			VariantGraph vg(arg(argc, argv, "-r", ""), arg(argc, argv, "-v", ""), prefix, o);
			printf("Num mutations: %lu num mutations-sample: %lu\nNum vars: %lu\n", vg.num_mutations, vg.num_mutations_samples, vg.num_vars);
			printf("Chromosome: %s #Vertices: %lu #Edges: %lu Seq length: %lu\n", vg.get_chr().c_str(), vg.get_num_vertices(), vg.get_num_edges(), vg.get_seq_length());
			vg.serialize();
			printf("Number of sample vector classes: %lu\n", vg.get_num_sample_classes());
			Index idx(&vg); idx.serialize(prefix);
			return 0;
		}
		if (cmd == "query") {
			std::string prefix = arg(argc, argv, "-p", "ser");
			Index idx(prefix);
			VariantGraph vg(prefix);
			printf("Chromosome: %s #Vertices: %lu #Edges: %lu Seq length: %lu\n", vg.get_chr().c_str(), vg.get_num_vertices(), vg.get_num_edges(), vg.get_seq_length());
			int type = atoi(arg(argc, argv, "-t", "6"));
			auto regions = read_regions(arg(argc, argv, "-r", "1"));
			std::string sample = arg(argc, argv, "-s", ""), outfile = arg(argc, argv, "-o", "");
			bool verbose = flag(argc, argv, "-v");
			auto t0 = std::chrono::steady_clock::now();
			for (size_t i = 0; i < regions.size(); i++) {
				if (type == 1) { std::vector<Variant> var; closest_var(&vg, &idx, regions[i].first, var, verbose, outfile); }
				else if (type == 2) query_sample_from_ref(&vg, &idx, regions[i].first, regions[i].second, sample, verbose, outfile);
				else if (type == 3) { bool hang = false; query_sample_from_sample(&vg, &idx, regions[i].first, regions[i].second, sample, verbose, outfile, nullptr, &hang); if (hang) { fprintf(stderr, "does not terminate\n"); return 3; } }
				else if (type == 5) { bool hang = false; get_sample_var_in_sample(&vg, &idx, regions[i].first, regions[i].second, sample, verbose, outfile, nullptr, nullptr, &hang); if (hang) { fprintf(stderr, "does not terminate\n"); return 3; } }
				else if (type == 4) get_sample_var_in_ref(&vg, &idx, regions[i].first, regions[i].second, sample, verbose, outfile);
				else if (type == 6) get_var_in_ref(&vg, &idx, regions[i].first, regions[i].second, verbose, outfile);
				else if (type == 7) {
					auto alts = read_sequences(arg(argc, argv, "-a", "")), refs = read_sequences(arg(argc, argv, "-b", ""));
					QueryLog log;
					auto s = samples_has_var(&vg, &idx, regions[i].first, refs.at(i), alts.at(i), verbose, outfile, &log);
					if (!log.err.empty()) fputs(log.err.c_str(), stderr);
				} else { fprintf(stderr, "Unsupported query type\n"); }
			}
			double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
			std::cout << "Query" << regions.size() << ": " << (type == 6 ? "(query_var_in_ref) " : "") << "Total Time Elapsed: " << std::to_string(dt) << "seconds" << std::endl;
			return 0;
		}
	} catch (const std::exception& e) { fprintf(stderr, "%s\n", e.what()); return 2; }
	return 1;
}
