// ORACLE — test infrastructure only (see vso.h).  Literal restatement of the query operators of
// include/query.h (graph walks, BFS objects, dedup, gate, stdout lines and -o file behaviour) and
// of read_regions / read_sequences from src/commands.cc.
#include "vso.h"

#include <algorithm>
#include <fstream>
#include <iostream>
#include <sstream>

namespace vso {

static const std::string REF = "ref";   // query.h:26

static void print_header(std::ostream& out) { out << "Pos\tRef\tAlt\tSamples\n"; }   // query.h:38-41
static void print_var(const Variant* var, std::ostream& out) {                           // query.h:43-50
	out << var->var_pos << "\t" << var->ref << "\t" << var->alt << "\t";
	for (const auto& sample : var->samples) out << sample.first << "(" << sample.second << ") ";
	out << std::endl;
}
static void dump_vars(const std::vector<Variant>& vars, const std::string& outfile) {
	std::ofstream out; out.open(outfile);
	print_header(out);
	for (const auto& var : vars) print_var(&var, out);
	out.close();
}
static void say(QueryLog* log, const std::string& line) { if (log) log->out += line; else std::cout << line; }
static void err(QueryLog* log, const std::string& line) { if (log) log->err += line + "\n"; }

Graph::vertex get_prev_vertex_with_sample(const VariantGraph* vg, const Index* idx, const uint64_t pos,
                                          const std::string& sample_id, uint64_t& ref_pos, uint64_t& sample_pos,
                                          bool* ub) {   // query.h:57-113
	uint64_t cur_pos = pos;
	uint64_t cur_ref_node_idx = 0;
	Graph::vertex v = idx->find(cur_pos, cur_ref_node_idx);
	Graph::vertex v_find = v;
	SampleInfo sample, sample_find;
	bool sample_found = false;
	while (true) {
		// the reference lets cur_ref_node_idx (uint64_t) wrap below zero and then indexes node_list out
		// of bounds (undefined behaviour); the oracle defines that case as "walk from vertex 0".
		if (cur_ref_node_idx > idx->node_list.size()) { if (ub) *ub = true; cur_ref_node_idx = 0; }
		v = idx->previous(cur_ref_node_idx);
		if (cur_ref_node_idx <= 1) {
			ref_pos = 1;
			v_find = v;
			vg->get_sample_from_vertex_if_exists(v_find, REF, sample_find);
			sample_pos = sample_find.index;
			break;
		}
		VariantGraph::BfsIterator it = vg->find(v, 1);
		++it;
		while (!it.done()) {
			v = (*it)->vertex_id;
			if (vg->get_sample_from_vertex_if_exists(v, REF, sample)) ref_pos = sample.index;
			if (vg->get_sample_from_vertex_if_exists(v, sample_id, sample_find)) {
				v_find = v; sample_found = true; sample_pos = sample_find.index;
			}
			++it;
			cur_ref_node_idx--;     // once per neighbour (query.h:103)
		}
		if (sample_found) break;
	}
	return v_find;
}

// ---- helpers shared by the sequence / sample-coordinate operators (t2, t3, t5) ----------------------

// out-neighbours of v in the order the reference's radius-1 BFS meets them (the iterator's first element is v itself)
static VariantGraph::BfsIterator neighbours_of(const VariantGraph* vg, Graph::vertex v) {
	VariantGraph::BfsIterator it = vg->find(v, 1);
	++it;
	return it;
}

// The four cutting rules both sequence operators apply to every vertex of the sample's path
// (query.h:157-173 in ref coordinates, :229-245 in sample coordinates): `at` = the running position on
// arrival, `next` = the position behind the vertex.  std::string::substr throws std::out_of_range
// exactly where the reference's calls do; the reference does not catch it (its process terminates),
// callers of the oracle do and report "threw".
struct SeqWindow {
	uint64_t lo, hi;
	bool open = false;             // record_seq
	std::string out;
	bool take(const std::string& vseq, uint64_t at, uint64_t next) {   // true: the walk is over
		const bool reaches_hi = next >= hi;
		if (open) {
			if (!reaches_hi) { out += vseq; return false; }
			out += vseq.substr(0, hi - at);
			return true;
		}
		if (next < lo) return false;
		if (!reaches_hi) { open = true; out += vseq.substr(lo - at); return false; }
		out = vseq.substr(lo - at, hi - lo);
		return true;
	}
};

static void write_sequence(const std::string& seq, const std::string& outfile) {   // :180-186, :252-258
	std::ofstream out; out.open(outfile); out << seq << std::endl; out.close();
}

// Start of t3 / t5 (query.h:201-214, :496-510): get_prev_vertex_with_sample from pos_x, repeated from
// the ref position of the vertex found while its sample position is still >= pos_x.  When that ref
// position maps to itself the reference never leaves the loop; the chain of positions is
// deterministic, so seeing one twice means exactly that (`hang`).
struct SampleStart { Graph::vertex v; uint64_t ref_pos = 0, sample_pos = 0; bool hang = false; };
static SampleStart start_before(const VariantGraph* vg, const Index* idx, uint64_t pos_x, const std::string& sample_id, bool* ub) {
	SampleStart st;
	st.v = get_prev_vertex_with_sample(vg, idx, pos_x, sample_id, st.ref_pos, st.sample_pos, ub);
	std::vector<uint64_t> tried;
	while (st.sample_pos >= pos_x && st.v > 0) {
		const uint64_t from = st.ref_pos;
		if (std::find(tried.begin(), tried.end(), from) != tried.end()) { st.hang = true; break; }
		tried.push_back(from);
		st.v = get_prev_vertex_with_sample(vg, idx, from, sample_id, st.ref_pos, st.sample_pos, ub);
	}
	return st;
}

// t2 — sample's sequence over ref positions [pos_x, pos_y) (query.h:120-189).  next_ref_pos of a vertex
// is the index of the FIRST ref-carrying out-neighbour (:143-151; t4 and t5 take the last one).
std::string query_sample_from_ref(const VariantGraph* vg, const Index* idx, const uint64_t pos_x, const uint64_t pos_y,
                                  const std::string& sample_id, bool print, const std::string& outfile, bool* ub) {
	uint64_t ref_pos = 0, sample_pos = 0;
	const Graph::vertex from = get_prev_vertex_with_sample(vg, idx, pos_x, sample_id, ref_pos, sample_pos, ub);
	SeqWindow w{pos_x, pos_y};
	for (VariantGraph::PathIterator it = vg->find(from, sample_id); !it.done(); ++it) {
		uint64_t next_ref_pos = ref_pos + (*it)->length;
		for (VariantGraph::BfsIterator nb = neighbours_of(vg, (*it)->vertex_id); !nb.done(); ++nb) {
			SampleInfo info;
			if (vg->get_sample_from_vertex_if_exists((*nb)->vertex_id, REF, info)) { next_ref_pos = info.index; break; }
		}
		if (w.take(vg->get_sequence(*(*it)), ref_pos, next_ref_pos)) break;
		ref_pos = next_ref_pos;
	}
	if (print) write_sequence(w.out, outfile);
	return w.out;
}

// t3 — the same over [pos_x, pos_y) of the sample's own coordinates (query.h:195-261): the running
// position starts at the sample's index in the start vertex and grows by the length of every vertex.
std::string query_sample_from_sample(const VariantGraph* vg, const Index* idx, const uint64_t pos_x, const uint64_t pos_y,
                                     const std::string& sample_id, bool print, const std::string& outfile, bool* ub, bool* hang) {
	const SampleStart st = start_before(vg, idx, pos_x, sample_id, ub);
	if (st.hang) { if (hang) *hang = true; return ""; }
	SeqWindow w{pos_x, pos_y};
	uint64_t sample_pos = st.sample_pos;
	for (VariantGraph::PathIterator it = vg->find(st.v, sample_id); !it.done(); ++it) {
		const uint64_t next_sample_pos = sample_pos + (*it)->length;
		if (w.take(vg->get_sequence(*(*it)), sample_pos, next_sample_pos)) break;
		sample_pos = next_sample_pos;
	}
	if (print) write_sequence(w.out, outfile);
	return w.out;
}

bool get_samples(const Vertex* v, const VariantGraph* vg, std::vector<std::pair<std::string, std::string>>& sample_ids) {   // :268-285
	bool is_var = false;
	sample_ids = {};
	for (size_t i = 0; i < v->s_info.size(); ++i) {
		std::string sample_id = vg->get_sample_name(vg->get_sample_id(*v, (uint32_t)i));
		if (sample_id != REF) {
			std::string phasing = vg->get_sample_phasing(*v, (uint32_t)i);
			sample_ids.push_back(std::make_pair(sample_id, phasing));
			is_var = true;
		}
	}
	return is_var;
}

bool next_variant_in_ref(const VariantGraph* vg, const Index* idx, const uint64_t pos, std::vector<Variant>& vars,
                         uint64_t& next_pos, const uint64_t end) {   // :297-436
	bool found_var = false;
	Graph::vertex v = idx->find(pos);
	VariantGraph::PathIterator it = vg->find(v, REF);
	VariantGraph::PathIterator next_it = vg->find(v, REF);
	++next_it;
	while (!it.done()) {
		SampleInfo ref_sample;
		if (vg->get_sample_from_vertex_if_exists((*it)->vertex_id, REF, ref_sample)) {
			if ((uint64_t)ref_sample.index + (*it)->length >= end) break;
		}
		VariantGraph::BfsIterator bfs_it = vg->find((*it)->vertex_id, 1);
		++bfs_it;
		while (!bfs_it.done()) {
			Variant var;
			if ((*bfs_it)->vertex_id == (*next_it)->vertex_id) { ++bfs_it; continue; }
			std::vector<std::pair<std::string, std::string>> sample_ids;
			if (get_samples((*bfs_it), vg, sample_ids)) {
				SampleInfo sample;
				if (vg->get_sample_from_vertex_if_exists((*bfs_it)->vertex_id, REF, sample)) {   // deletion :336-350
					var.ref = vg->get_sequence(*(*next_it));
					var.alt = "";
					var.samples = sample_ids;
					if (vg->get_sample_from_vertex_if_exists((*next_it)->vertex_id, REF, sample)) { var.var_pos = sample.index; var.var_pos_set = true; }
				} else {
					vg->get_sample_from_vertex_if_exists((*it)->vertex_id, REF, sample);
					uint64_t prev_ref_idx = sample.index;
					VariantGraph::PathIterator dfs_it = vg->find((*bfs_it)->vertex_id, sample_ids[0].first);
					++dfs_it;
					vg->get_sample_from_vertex_if_exists((*dfs_it)->vertex_id, REF, sample);   // on failure `sample` keeps its value (:359-365)
					uint64_t next_ref_idx = sample.index;
					std::string prev_ref = vg->get_sequence(*(*it));
					if (next_ref_idx == prev_ref_idx + prev_ref.length()) {        // insertion :369-376
						var.ref = "";
						var.alt = vg->get_sequence(*(*bfs_it));
						var.samples = sample_ids;
						var.var_pos = next_ref_idx - 1; var.var_pos_set = true;
					} else {                                                     // substitution :378-392
						var.alt = vg->get_sequence(*(*bfs_it));
						var.ref = vg->get_sequence(*(*next_it));
						var.samples = sample_ids;
						if (vg->get_sample_from_vertex_if_exists((*next_it)->vertex_id, REF, sample)) { var.var_pos = sample.index; var.var_pos_set = true; }
					}
				}
			}
			// only add var if not seen before (:397-414)
			if (vars.size() < 1 || (vars.back().var_pos != var.var_pos || vars.back().alt != var.alt)) {
				bool found_same = false;
				if (vars.size() > 1 && vars.back().var_pos == var.var_pos) {
					for (auto rit = vars.rbegin(); rit != vars.rend(); ++rit) {
						if (rit->var_pos < var.var_pos) break;
						if (rit->var_pos == var.var_pos && rit->alt == var.alt) { found_same = true; break; }
					}
				}
				if (!found_same) { found_var = true; vars.push_back(var); }
			}
			++bfs_it;
		}
		if (found_var) break;
		++it;
		++next_it;
	}
	SampleInfo sample;
	if (vg->get_sample_from_vertex_if_exists((*next_it)->vertex_id, REF, sample)) next_pos = sample.index;
	return found_var;
}

bool closest_var(const VariantGraph* vg, const Index* idx, const uint64_t pos, std::vector<Variant>& vars,
                 bool print, const std::string& outfile) {   // :441-483
	std::vector<Variant> next_var;
	uint64_t next_pos;
	if (next_variant_in_ref(vg, idx, pos, next_var, next_pos)) {
		uint64_t next_var_pos = next_var[0].var_pos;
		std::vector<Variant> prev_var;
		int cur_pos = (int)(pos - (next_var_pos - pos));
		if (cur_pos > 0) {
			next_variant_in_ref(vg, idx, cur_pos, prev_var, next_pos);
			// the reference reads prev_var[0] even when that call found nothing (possible when the next
			// variant's pos is below `pos`, so the mirrored position lies beyond it): defined as "keep next_var"
			if (!prev_var.empty() && prev_var[0].var_pos != next_var_pos) vars = prev_var; else vars = next_var;
		} else vars = next_var;
	} else {
		int cur_pos = (int)pos - 1;
		while (cur_pos > 0 && !next_variant_in_ref(vg, idx, cur_pos, next_var, next_pos)) {
			if (cur_pos == 1) return false;
			cur_pos--;
		}
		vars = next_var;
	}
	if (print) dump_vars(vars, outfile);
	return true;
}

// t5 — a sample's variants over [pos_x, pos_y) of its own coordinates (query.h:490-612).  Start as t3
// (incl. the loop that may never end), then from the backbone vertex holding ref_pos (:513) along the
// sample's path: a vertex carrying the sample is a row when pos_x < sample_pos (strictly) on arrival,
// the walk stops at sample_pos >= pos_y.  next_ref_pos / next_ref come from the LAST ref-carrying
// out-neighbour (:541-549).
std::vector<Variant> get_sample_var_in_sample(const VariantGraph* vg, const Index* idx, const uint64_t pos_x, const uint64_t pos_y,
                                              const std::string& sample_id, bool print, const std::string& outfile, QueryLog* log,
                                              bool* ub, bool* hang) {
	std::vector<Variant> rows;
	const SampleStart st = start_before(vg, idx, pos_x, sample_id, ub);
	if (st.hang) { if (hang) *hang = true; return rows; }
	uint64_t ref_pos = st.ref_pos, sample_pos = st.sample_pos;
	const Graph::vertex first = idx->find(ref_pos);
	SampleInfo info;
	if (vg->get_sample_from_vertex_if_exists(first, REF, info)) {       // :515-522: back to the start of that vertex
		sample_pos -= ref_pos - info.index;
		ref_pos = info.index;
	} else err(log, "reference node is expected to be found!");
	std::string ref_before;                                            // cur_ref
	for (VariantGraph::PathIterator it = vg->find(first, sample_id); !it.done(); ++it) {
		if (sample_pos >= pos_y) break;
		const Vertex* cur = *it;
		uint64_t next_ref_pos = ref_pos + cur->length;
		std::string next_ref;
		for (VariantGraph::BfsIterator nb = neighbours_of(vg, cur->vertex_id); !nb.done(); ++nb)
			if (vg->get_sample_from_vertex_if_exists((*nb)->vertex_id, REF, info)) { next_ref_pos = info.index; next_ref = vg->get_sequence(*(*nb)); }
		SampleInfo mine;
		if (sample_pos > pos_x && vg->get_sample_from_vertex_if_exists(cur->vertex_id, sample_id, mine)) {
			Variant row;
			row.var_pos_set = true;
			if (ref_pos == next_ref_pos) {                                   // insertion :556-562
				ref_before = "";
				row.alt = vg->get_sequence(*cur);
				row.var_pos = ref_pos;
			} else if (vg->get_sample_from_vertex_if_exists(cur->vertex_id, REF, info)) {   // deletion :564-573
				row.var_pos = mine.index;
				ref_before = vg->get_sequence(vg->get_vertex(idx->find(ref_pos - 1)));
			} else {                                                         // substitution :574-579
				row.alt = vg->get_sequence(*cur);
				row.var_pos = mine.index;
			}
			row.ref = ref_before;
			get_samples(cur, vg, row.samples);
			rows.push_back(row);
		}
		ref_before = next_ref;
		ref_pos = next_ref_pos;
		sample_pos += cur->length;
	}
	say(log, "Number of variants get_sample_var_in_sample: " + std::to_string(rows.size()) + "\n");
	if (print) dump_vars(rows, outfile);
	return rows;
}

std::vector<Variant> get_sample_var_in_ref(const VariantGraph* vg, const Index* idx, const uint64_t pos_x,
                                           const uint64_t pos_y, const std::string& sample_id, bool print,
                                           const std::string& outfile, QueryLog* log, bool* ub) {   // :618-729
	std::vector<Variant> vars;
	uint64_t ref_pos = 0, sample_pos = 0;
	if (idx->is_empty(pos_x, pos_y)) {
		say(log, "Number of variants get_sample_var_in_ref: " + std::to_string(vars.size()) + "\n");
		if (print) dump_vars(vars, outfile);
		return vars;
	}
	Graph::vertex closest_v = get_prev_vertex_with_sample(vg, idx, pos_x, sample_id, ref_pos, sample_pos, ub);
	SampleInfo sample;
	VariantGraph::PathIterator it = vg->find(closest_v, sample_id);
	std::string cur_ref;
	while (!it.done()) {
		if (ref_pos >= pos_y) break;
		Graph::vertex cur_v = (*it)->vertex_id;
		Variant var;
		uint64_t l = (*it)->length;
		uint64_t next_ref_pos = ref_pos + l;
		std::string next_ref;
		VariantGraph::BfsIterator bfs_it = vg->find((*it)->vertex_id, 1);
		++bfs_it;
		while (!bfs_it.done()) {          // last ref-carrying neighbour wins (:667-674)
			Graph::vertex v = (*bfs_it)->vertex_id;
			if (vg->get_sample_from_vertex_if_exists(v, REF, sample)) {
				next_ref_pos = sample.index;
				next_ref = vg->get_sequence(*(*bfs_it));
			}
			++bfs_it;
		}
		if (ref_pos >= pos_x && vg->get_sample_from_vertex_if_exists(cur_v, sample_id, sample)) {
			std::string alt;
			if (ref_pos == next_ref_pos) {                                   // insertion :682-688
				cur_ref = "";
				alt = vg->get_sequence(*(*it));
				var.var_pos = ref_pos - 1; var.var_pos_set = true;
			} else if (vg->get_sample_from_vertex_if_exists(cur_v, REF, sample)) {   // deletion :690-698
				alt = "";
				Graph::vertex v = idx->find(ref_pos - 1);
				cur_ref = vg->get_sequence(vg->get_vertex(v));
				if (vg->get_sample_from_vertex_if_exists(v, REF, sample)) { var.var_pos = sample.index; var.var_pos_set = true; }
			} else {                                                         // substitution :699-703
				alt = vg->get_sequence(*(*it));
				var.var_pos = ref_pos; var.var_pos_set = true;
			}
			var.alt = alt;
			var.ref = cur_ref;
			get_samples((*it), vg, var.samples);
			vars.push_back(var);
		}
		cur_ref = next_ref;
		ref_pos = next_ref_pos;
		++it;
	}
	say(log, "Number of variants get_sample_var_in_ref: " + std::to_string(vars.size()) + "\n");
	if (print) dump_vars(vars, outfile);
	return vars;
}

std::vector<Variant> get_var_in_ref(const VariantGraph* vg, const Index* idx, const uint64_t pos_x, const uint64_t pos_y,
                                    bool print, const std::string& outfile, QueryLog* log) {   // :736-784
	std::vector<Variant> vars;
	if (idx->is_empty(pos_x, pos_y)) {
		say(log, "Number of variants get_sample_var_in_ref: " + std::to_string(vars.size()) + "\n");   // sic (:746)
		if (print) dump_vars(vars, outfile);
		return vars;
	}
	uint64_t cur_pos = pos_x;
	while (cur_pos < pos_y) {
		uint64_t next_pos = 0;
		if (next_variant_in_ref(vg, idx, cur_pos, vars, next_pos, pos_y)) {
			cur_pos = next_pos;
			if (cur_pos >= pos_y) break;
		} else break;
	}
	say(log, "Number of variants get_var_in_ref: " + std::to_string(vars.size()) + "\n");
	if (print) dump_vars(vars, outfile);
	return vars;
}

std::vector<std::pair<std::string, std::string>>
samples_has_var(const VariantGraph* vg, const Index* idx, const uint64_t pos, const std::string& ref,
                const std::string& alt, bool print, const std::string& outfile, QueryLog* log) {   // :792-823
	std::vector<Variant> vars;
	uint64_t next_pos = 0;
	std::vector<std::pair<std::string, std::string>> samples;
	next_variant_in_ref(vg, idx, pos, vars, next_pos);
	for (const auto& var : vars) {
		if (var.ref == ref && var.var_pos == pos && var.alt == alt) {
			samples.assign(var.samples.begin(), var.samples.end());
			if (print) {
				std::ofstream out; out.open(outfile);
				for (auto i = samples.begin(); i != samples.end(); ++i) out << i->first << ' ' << i->second;
				out << std::endl;
				out.close();
			}
			return samples;
		}
	}
	err(log, "There is no such variant!");
	return samples;
}

std::vector<std::pair<uint64_t, uint64_t>> read_regions(std::string region) {   // commands.cc:64-93
	std::vector<std::pair<uint64_t, uint64_t>> regions;
	auto pos = region.find(',');
	while (true) {
		std::string token = region.substr(0, pos);
		auto pos2 = token.find(':');
		uint64_t beg = 0, end = 0;
		if (pos2 == std::string::npos) beg = std::stoi(token);
		else { end = std::stoi(token.substr(pos2 + 1)); beg = std::stoi(token.substr(0, pos2)); }
		regions.push_back(std::make_pair(beg, end));
		if (pos == std::string::npos) break;
		region = region.substr(pos + 1);
		pos = region.find(',');
	}
	std::sort(regions.begin(), regions.end());
	return regions;
}

std::vector<std::string> read_sequences(std::string s) {   // commands.cc:96-111
	std::vector<std::string> seqs;
	auto pos = s.find(',');
	while (true) {
		seqs.push_back(s.substr(0, pos));
		if (pos == std::string::npos) break;
		s = s.substr(pos + 1);
		pos = s.find(',');
	}
	return seqs;
}

}  // namespace vso
