// vsgpu_query — front-end with the reference's `variantstore query` flags
// (src/variantstore.cc:101-134, src/commands.cc:113-215):
//   -p <ser prefix> -t <1..7> -r <beg[:end][,...]> -m <0|1> [-o file] [-s sample] [-a alts] [-b refs] [-v]
// -m is accepted and ignored (both modes give identical results; only the reference's paging differs).
// Beyond the reference's flags, for batches that do not fit a command line and for more than one contig:
//   --regions-file <file>   one region per line, `beg[:end]` (same parse and the same sort as -r: commands.cc:64-93);
//                           with --prefixes every line is `<contig> beg[:end]`
//   --prefixes <p1,p2,...>  several ser/ directories (contigs, or position ranges of one: `prefix@lo-hi`) behind the router
//                           (types 4 and 6); --devices <n> GPUs (default: all)
// Unlike query_main's per-region loop the whole region list goes to the GPU as one batch; the lines
// printed per region are the reference's.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <stdexcept>
#include <tuple>
#include <vector>

#include "../../include/vsgpu.h"

static std::vector<std::tuple<uint64_t, uint64_t>> read_regions(std::string region) {   // commands.cc:64-93
	std::vector<std::tuple<uint64_t, uint64_t>> regions;
	auto pos = region.find(',');
	while (true) {
		std::string token = region.substr(0, pos);
		auto pos2 = token.find(':');
		uint64_t beg = 0, end = 0;
		if (pos2 == std::string::npos) beg = std::stoi(token);
		else { end = std::stoi(token.substr(pos2 + 1)); beg = std::stoi(token.substr(0, pos2)); }
		regions.push_back(std::make_tuple(beg, end));
		if (pos == std::string::npos) break;
		region = region.substr(pos + 1);
		pos = region.find(',');
	}
	std::sort(regions.begin(), regions.end());
	return regions;
}
static void parse_region_token(const std::string& token, uint64_t& beg, uint64_t& end) {   // one `beg[:end]` as read_regions parses it
	auto pos2 = token.find(':');
	beg = 0; end = 0;
	if (pos2 == std::string::npos) beg = std::stoi(token);
	else { end = std::stoi(token.substr(pos2 + 1)); beg = std::stoi(token.substr(0, pos2)); }
}
// --regions-file: the region list of -r, one region per line (blank lines and lines starting with '#' skipped), sorted as read_regions sorts
static std::vector<std::tuple<uint64_t, uint64_t>> read_regions_file(const char* path) {
	std::vector<std::tuple<uint64_t, uint64_t>> regions;
	std::ifstream f(path);
	if (!f.good()) throw std::runtime_error(std::string("cannot open ") + path);
	std::string line;
	while (std::getline(f, line)) {
		while (!line.empty() && (line.back() == '\r' || line.back() == ' ' || line.back() == '\t')) line.pop_back();
		if (line.empty() || line[0] == '#') continue;
		uint64_t b, e; parse_region_token(line, b, e);
		regions.push_back(std::make_tuple(b, e));
	}
	std::sort(regions.begin(), regions.end());
	return regions;
}
static std::vector<std::string> read_sequences(std::string s) {   // commands.cc:96-111
	std::vector<std::string> seqs;
	auto pos = s.find(',');
	while (true) { seqs.push_back(s.substr(0, pos)); if (pos == std::string::npos) break; s = s.substr(pos + 1); pos = s.find(','); }
	return seqs;
}
static const char* opt(int argc, char** argv, const char* f, const char* d) { for (int i = 1; i + 1 < argc; i++) if (!strcmp(argv[i], f)) return argv[i + 1]; return d; }
static bool flag(int argc, char** argv, const char* f) { for (int i = 1; i < argc; i++) if (!strcmp(argv[i], f)) return true; return false; }
static void write_rows(const std::string& outfile, const char* text, bool header) {
	std::ofstream out; out.open(outfile);                 // truncating rewrite per call, as the reference does
	if (header) out << "Pos\tRef\tAlt\tSamples\n";
	out << text; out.close();
}

// --prefixes: several contigs (or position ranges of one) on the GPUs of this node through vsgpu_router_*; types 4 and 6.
// Regions come from --regions-file as `<contig> beg[:end]` lines and are answered in (contig of the router's table, beg, end)
// order — per contig the order the reference's read_regions gives its single contig.
static int router_main(int argc, char** argv, const char* prefixes, int type, const char* rfile) {
	if (type != 4 && type != 6) { fprintf(stderr, "--prefixes serves query types 4 and 6\n"); return 1; }
	if (!rfile) { fprintf(stderr, "--prefixes needs --regions-file with `<contig> beg[:end]` lines\n"); return 1; }
	std::vector<std::string> pre; std::vector<uint64_t> lo, hi;
	for (const std::string& tok : read_sequences(prefixes)) {
		auto at = tok.rfind('@');
		uint64_t a = 0, b = 0; std::string p = tok;
		if (at != std::string::npos && tok.find('-', at) != std::string::npos) { p = tok.substr(0, at); a = std::stoull(tok.substr(at + 1)); b = std::stoull(tok.substr(tok.find('-', at) + 1)); }
		pre.push_back(p); lo.push_back(a); hi.push_back(b);
	}
	std::vector<const char*> pp; for (auto& p : pre) pp.push_back(p.c_str());
	vsgpu_router* r = nullptr;
	if (vsgpu_router_open((uint32_t)pp.size(), pp.data(), lo.data(), hi.data(), nullptr, atoi(opt(argc, argv, "--devices", "0")), &r) != 0) { fprintf(stderr, "%s\n", vsgpu_router_last_error()); return 2; }
	const std::string sample = opt(argc, argv, "-s", "");
	std::vector<std::tuple<uint32_t, uint64_t, uint64_t>> regions;
	{
		std::ifstream f(rfile);
		if (!f.good()) { fprintf(stderr, "cannot open %s\n", rfile); vsgpu_router_close(r); return 1; }
		std::string line;
		while (std::getline(f, line)) {
			while (!line.empty() && (line.back() == '\r' || line.back() == ' ' || line.back() == '\t')) line.pop_back();
			if (line.empty() || line[0] == '#') continue;
			auto sp = line.find_first_of(" \t");
			if (sp == std::string::npos) { fprintf(stderr, "bad line (want `<contig> beg[:end]`): %s\n", line.c_str()); vsgpu_router_close(r); return 1; }
			uint32_t c = 0;
			if (vsgpu_router_contig_id(r, line.substr(0, sp).c_str(), &c) != 0) { fprintf(stderr, "%s\n", vsgpu_router_last_error()); vsgpu_router_close(r); return 2; }
			uint64_t b, e;
			try { parse_region_token(line.substr(line.find_first_not_of(" \t", sp)), b, e); } catch (const std::exception&) { fprintf(stderr, "bad region: %s\n", line.c_str()); vsgpu_router_close(r); return 1; }
			regions.push_back(std::make_tuple(c, b, e));
		}
	}
	std::sort(regions.begin(), regions.end());
	const uint64_t n = regions.size();
	std::vector<uint32_t> c(n), x(n), y(n), s(n, 0), shard(n), rlo(n), c6(n), c4(n);
	for (uint64_t i = 0; i < n; i++) { c[i] = std::get<0>(regions[i]); x[i] = (uint32_t)std::get<1>(regions[i]); y[i] = (uint32_t)std::get<2>(regions[i]); }
	// the sample id is looked up per shard (every ser/ has its own sampleid_map.lst); t6 does not need one
	for (uint64_t i = 0; i < n; i++) s[i] = 1;
	if (type == 4) {
		// route once to learn the shards, then set each region's sample id from its shard's map
		std::vector<uint32_t> sid(vsgpu_router_num_shards(r), 0);
		for (uint32_t k = 0; k < sid.size(); k++) if (vsgpu_sample_id(vsgpu_router_shard_index(r, k), sample.c_str(), &sid[k]) != 0) { fprintf(stderr, "%s\n", vsgpu_last_error()); vsgpu_router_close(r); return 2; }
		if (vsgpu_router_query_t6t4(r, n, c.data(), x.data(), y.data(), s.data(), shard.data(), rlo.data(), c6.data(), c4.data()) != 0) { fprintf(stderr, "%s\n", vsgpu_router_last_error()); vsgpu_router_close(r); return 2; }
		bool same = true;
		for (uint64_t i = 0; i < n; i++) { if (sid[shard[i]] != s[i]) same = false; s[i] = sid[shard[i]]; }
		if (same) goto print;
	}
	if (vsgpu_router_query_t6t4(r, n, c.data(), x.data(), y.data(), s.data(), shard.data(), rlo.data(), c6.data(), c4.data()) != 0) { fprintf(stderr, "%s\n", vsgpu_router_last_error()); vsgpu_router_close(r); return 2; }
print:
	for (uint64_t i = 0; i < n; i++) {
		if (type == 6) printf("%s\tNumber of variants %s: %u\n", vsgpu_router_contig_name(r, c[i]), rlo[i] == VSGPU_NONE ? "get_sample_var_in_ref" : "get_var_in_ref", c6[i]);
		else printf("%s\tNumber of variants get_sample_var_in_ref: %u\n", vsgpu_router_contig_name(r, c[i]), c4[i]);
	}
	vsgpu_router_close(r);
	return 0;
}

int main(int argc, char** argv) {
	int a0 = (argc > 1 && !strcmp(argv[1], "query")) ? 2 : 1;
	argc -= a0 - 1; argv += a0 - 1;
	const char* prefix = opt(argc, argv, "-p", nullptr); const char* tstr = opt(argc, argv, "-t", nullptr); const char* rstr = opt(argc, argv, "-r", nullptr);
	const char* rfile = opt(argc, argv, "--regions-file", nullptr); const char* prefixes = opt(argc, argv, "--prefixes", nullptr);
	if ((!prefix && !prefixes) || !tstr || (!rstr && !rfile)) {
		fprintf(stderr, "usage: vsgpu_query [query] -p <prefix> -t <1..7> -r <regions> -m <0|1> [-o file] [-s sample] [-a alts] [-b refs] [-v]\n"
		                "       [--regions-file <file>] [--prefixes <p1,p2,...> [--devices <n>]]\n");
		return 1;
	}
	const int type = atoi(tstr);
	if (prefixes) return router_main(argc, argv, prefixes, type, rfile);
	const std::string outfile = opt(argc, argv, "-o", ""), sample = opt(argc, argv, "-s", "");
	const bool verbose = flag(argc, argv, "-v");
	vsgpu_index* idx = nullptr;
	if (vsgpu_open(prefix, atoi(opt(argc, argv, "--device", "0")), &idx) != 0) { fprintf(stderr, "%s\n", vsgpu_last_error()); return 2; }
	vsgpu_info_t info; vsgpu_info(idx, &info);
	printf("Chromosome: %s #Vertices: %lu #Edges: 0 Seq length: %lu\n", info.chr, (unsigned long)info.num_vertices_cqf, (unsigned long)info.seq_length);
	std::vector<std::tuple<uint64_t, uint64_t>> regions;
	try { regions = rfile ? read_regions_file(rfile) : read_regions(rstr); } catch (const std::exception& e) { fprintf(stderr, "%s\n", e.what()); vsgpu_close(idx); return 1; }
	const uint64_t n = regions.size();
	std::vector<uint64_t> x(n), y(n);
	for (uint64_t i = 0; i < n; i++) { x[i] = std::get<0>(regions[i]); y[i] = std::get<1>(regions[i]); }
	auto t0 = std::chrono::steady_clock::now();
	int rc = 0;
	if (type == 6) {
		std::vector<uint32_t> lo(n), hi(n), cnt(n);
		rc = vsgpu_query_t6(idx, n, x.data(), y.data(), lo.data(), hi.data(), cnt.data());
		for (uint64_t i = 0; i < n && rc == 0; i++) {
			// the reference prints the t4 label when its is_empty gate fires (query.h:746)
			printf("Number of variants %s: %u\n", lo[i] == VSGPU_NONE ? "get_sample_var_in_ref" : "get_var_in_ref", cnt[i]);
		}
		if (verbose && rc == 0 && n) {
			// every call of the reference truncates and rewrites -o (query.h:774-781), so what is left on disk
			// are the rows of the last region: rendered on the device, written once
			vsgpu_text* text = nullptr;
			rc = vsgpu_render_t6(idx, 1, &x[n - 1], &y[n - 1], 1, &text);
			if (rc == 0) { write_rows(outfile, vsgpu_text_bytes(text), true); vsgpu_text_free(text); }
		}
	} else if (type == 4) {
		uint32_t sid = 0;
		if (vsgpu_sample_id(idx, sample.c_str(), &sid) != 0) { fprintf(stderr, "%s\n", vsgpu_last_error()); return 2; }
		std::vector<uint32_t> s(n, sid);
		vsgpu_result* res = nullptr;
		rc = vsgpu_query_t4(idx, n, x.data(), y.data(), s.data(), &res);
		if (rc == 0) {
			const uint64_t* off = vsgpu_result_offsets(res);
			for (uint64_t i = 0; i < n; i++) printf("Number of variants get_sample_var_in_ref: %lu\n", (unsigned long)(off[i + 1] - off[i]));
			vsgpu_result_free(res);
			if (verbose && n) {                                    // as for type 6: -o ends up holding the last region's rows (query.h:719-726)
				vsgpu_text* text = nullptr;
				rc = vsgpu_render_t4(idx, 1, &x[n - 1], &y[n - 1], &sid, 1, &text);
				if (rc == 0) { write_rows(outfile, vsgpu_text_bytes(text), true); vsgpu_text_free(text); }
			}
		}
	} else if (type == 1) {
		// closest_var prints nothing; with -v every call that finds a variant rewrites -o (query.h:472-479)
		std::vector<uint32_t> lo(n), hi(n);
		rc = vsgpu_query_t1(idx, n, x.data(), lo.data(), hi.data());
		for (uint64_t i = 0; i < n && rc == 0 && verbose; i++) {
			if (lo[i] == VSGPU_NONE) continue;                     // the operator returned false before printing
			char* text = nullptr; uint64_t nrows = 0;
			if (vsgpu_rows_t1(idx, lo[i], hi[i], 1, &text, &nrows) == 0) { write_rows(outfile, text, true); vsgpu_free(text); }
		}
	} else if (type == 5) {
		uint32_t sid = 0;
		if (vsgpu_sample_id(idx, sample.c_str(), &sid) != 0) { fprintf(stderr, "%s\n", vsgpu_last_error()); return 2; }
		std::vector<uint32_t> s(n, sid);
		vsgpu_result* res = nullptr;
		rc = vsgpu_query_t5(idx, n, x.data(), y.data(), s.data(), &res);
		if (rc == 0) {
			const uint64_t* off = vsgpu_result_offsets(res); const uint8_t* st = vsgpu_result_status(res);
			for (uint64_t i = 0; i < n; i++) {
				if (st[i] == 2) { fprintf(stderr, "region %lu:%lu: the reference never returns from get_sample_var_in_sample (query.h:505-510)\n", (unsigned long)x[i], (unsigned long)y[i]); vsgpu_result_free(res); vsgpu_close(idx); return 3; }
				printf("Number of variants get_sample_var_in_sample: %lu\n", (unsigned long)(off[i + 1] - off[i]));
			}
			vsgpu_result_free(res);
			if (verbose && n) {                                    // -o ends up holding the last region's rows (query.h:596-606)
				vsgpu_text* text = nullptr;
				rc = vsgpu_render_t5(idx, 1, &x[n - 1], &y[n - 1], &sid, 1, &text);
				if (rc == 0) { write_rows(outfile, vsgpu_text_bytes(text), true); vsgpu_text_free(text); }
			}
		}
	} else if (type == 2 || type == 3) {
		// query_sample_from_ref / query_sample_from_sample print nothing; with -v every call rewrites -o with "<sequence>\n"
		// (query.h:180-186), so the file ends up holding the last region's.  Where the reference's substr
		// throws std::out_of_range its process terminates there: same message, same kind of exit.
		uint32_t sid = 0;
		if (vsgpu_sample_id(idx, sample.c_str(), &sid) != 0) { fprintf(stderr, "%s\n", vsgpu_last_error()); return 2; }
		std::vector<uint32_t> s(n, sid);
		vsgpu_text* text = nullptr;
		rc = type == 2 ? vsgpu_query_t2(idx, n, x.data(), y.data(), s.data(), &text) : vsgpu_query_t3(idx, n, x.data(), y.data(), s.data(), &text);
		if (rc == 0) {
			const uint64_t* off = vsgpu_text_offsets(text); const char* bytes = vsgpu_text_bytes(text); const uint8_t* st = vsgpu_text_status(text);
			for (uint64_t i = 0; i < n; i++) {
				if (st[i] == 2) { fprintf(stderr, "region %lu:%lu: the reference never returns from query_sample_from_sample (query.h:209-214)\n", (unsigned long)x[i], (unsigned long)y[i]); vsgpu_text_free(text); vsgpu_close(idx); return 3; }
				if (st[i]) { fprintf(stderr, "terminate called after throwing an instance of 'std::out_of_range'\n  what():  basic_string::substr: __pos > this->size()\n"); vsgpu_text_free(text); vsgpu_close(idx); return 134; }
				if (verbose) { std::string t(bytes + off[i], bytes + off[i + 1]); t += "\n"; write_rows(outfile, t.c_str(), false); }
			}
			vsgpu_text_free(text);
		}
	} else if (type == 7) {
		auto alts = read_sequences(opt(argc, argv, "-a", "")), refs = read_sequences(opt(argc, argv, "-b", ""));
		if (alts.size() < n || refs.size() < n) { fprintf(stderr, "-a/-b need one entry per position\n"); return 1; }
		std::vector<const char*> rp(n), ap(n);
		for (uint64_t i = 0; i < n; i++) { rp[i] = refs[i].c_str(); ap[i] = alts[i].c_str(); }
		std::vector<uint32_t> rec(n);
		rc = vsgpu_query_t7(idx, n, x.data(), rp.data(), ap.data(), rec.data());
		for (uint64_t i = 0; i < n && rc == 0; i++) {
			if (rec[i] == VSGPU_NONE) { fprintf(stderr, "There is no such variant!\n"); continue; }
			if (verbose) { char* text = nullptr; uint64_t nc; if (vsgpu_rows_t7(idx, rec[i], &text, &nc) == 0) { std::string t = std::string(text) + "\n"; write_rows(outfile, t.c_str(), false); vsgpu_free(text); } }
		}
	} else { fprintf(stderr, "Unsupported query type\n"); rc = 1; }
	if (rc != 0) { fprintf(stderr, "%s\n", vsgpu_last_error()); vsgpu_close(idx); return 2; }
	double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	std::cout << "Query" << n << ": " << (type == 6 ? "(query_var_in_ref) " : "") << "Total Time Elapsed: " << std::to_string(dt) << "seconds" << std::endl;
	vsgpu_close(idx);
	return 0;
}
