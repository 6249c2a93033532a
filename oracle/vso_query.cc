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

// Sample's sequence in ref coordinates [pos_x, pos_y) (query.h:120-189).  std::string::substr throws
// std::out_of_range exactly where the reference's does (:163, :167) — the reference does not catch it,
// so its process terminates; callers of the oracle catch it and report "threw".
std::string query_sample_from_ref(const VariantGraph* vg, const Index* idx, const uint64_t pos_x, const uint64_t pos_y,
                                  const std::string& sample_id, bool print, const std::string& outfile, bool* ub) {
	std::string seq = "";
	uint64_t ref_pos = 0, sample_pos = 0;
	Graph::vertex closest_v = get_prev_vertex_with_sample(vg, idx, pos_x, sample_id, ref_pos, sample_pos, ub);
	VariantGraph::PathIterator it = vg->find(closest_v, sample_id);
	bool record_seq = false;
	std::string temp;
	while (!it.done()) {
		temp.assign(vg->get_sequence(*(*it)));
		uint64_t l = (*it)->length;
		uint64_t next_ref_pos = ref_pos + l;
		VariantGraph::BfsIterator bfs_it = vg->find((*it)->vertex_id, 1);
		++bfs_it;
		while (!bfs_it.done()) {          // FIRST ref-carrying neighbour wins (:143-151) — t4 takes the last
			Graph::vertex v = (*bfs_it)->vertex_id;
			SampleInfo sample;
			if (vg->get_sample_from_vertex_if_exists(v, REF, sample)) { next_ref_pos = sample.index; break; }
			++bfs_it;
		}
		if (record_seq == true && next_ref_pos < pos_y) {
			seq += temp;
		} else if (record_seq == true && next_ref_pos >= pos_y) {
			seq += temp.substr(0, pos_y - ref_pos);
			break;
		} else if (next_ref_pos >= pos_x && next_ref_pos < pos_y) {
			record_seq = true;
			seq += temp.substr(pos_x - ref_pos);
		} else if (next_ref_pos >= pos_x && next_ref_pos >= pos_y) {
			seq = temp.substr(pos_x - ref_pos, pos_y - pos_x);
			break;
		}
		++it;
		ref_pos = next_ref_pos;
	}
	if (print) { std::ofstream out; out.open(outfile); out << seq << std::endl; out.close(); }
	return seq;
}

// Sample's sequence in the sample's own coordinates [pos_x, pos_y) (query.h:195-261).  The loop at
// :209-214 repeats get_prev_vertex_with_sample from the ref position of the vertex found; when that
// position maps to itself the reference never leaves the loop — detected here (the chain of positions
// is deterministic, so revisiting one means it cycles) and reported through *hang.
std::string query_sample_from_sample(const VariantGraph* vg, const Index* idx, const uint64_t pos_x, const uint64_t pos_y,
                                     const std::string& sample_id, bool print, const std::string& outfile, bool* ub, bool* hang) {
	std::string seq = "";
	uint64_t ref_pos = 0, sample_pos = 0;
	Graph::vertex closest_v = get_prev_vertex_with_sample(vg, idx, pos_x, sample_id, ref_pos, sample_pos, ub);
	std::vector<uint64_t> seen;
	while (sample_pos >= pos_x && closest_v > 0) {
		uint64_t pos = ref_pos;
		if (std::find(seen.begin(), seen.end(), pos) != seen.end()) { if (hang) *hang = true; return ""; }
		seen.push_back(pos);
		closest_v = get_prev_vertex_with_sample(vg, idx, pos, sample_id, ref_pos, sample_pos, ub);
	}
	VariantGraph::PathIterator it = vg->find(closest_v, sample_id);
	bool record_seq = false;
	std::string temp;
	while (!it.done()) {
		temp.assign(vg->get_sequence(*(*it)));
		uint64_t l = (*it)->length;
		uint64_t next_sample_pos = sample_pos + l;
		if (record_seq == true && next_sample_pos < pos_y) {
			seq += temp;
		} else if (record_seq == true && next_sample_pos >= pos_y) {
			seq += temp.substr(0, pos_y - sample_pos);
			break;
		} else if (next_sample_pos >= pos_x && next_sample_pos < pos_y) {
			record_seq = true;
			seq += temp.substr(pos_x - sample_pos);
		} else if (next_sample_pos >= pos_x && next_sample_pos >= pos_y) {
			seq = temp.substr(pos_x - sample_pos, pos_y - pos_x);
			break;
		}
		++it;
		sample_pos = next_sample_pos;
	}
	if (print) { std::ofstream out; out.open(outfile); out << seq << std::endl; out.close(); }
	return seq;
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

// All variants of a sample over [pos_x, pos_y) of the sample's own coordinates (query.h:490-612).  Same
// start as query_sample_from_sample, incl. the loop that may never end (*hang).
std::vector<Variant> get_sample_var_in_sample(const VariantGraph* vg, const Index* idx, const uint64_t pos_x, const uint64_t pos_y,
                                              const std::string& sample_id, bool print, const std::string& outfile, QueryLog* log,
                                              bool* ub, bool* hang) {
	std::vector<Variant> vars;
	uint64_t ref_pos = 0, sample_pos = 0;
	Graph::vertex closest_v = get_prev_vertex_with_sample(vg, idx, pos_x, sample_id, ref_pos, sample_pos, ub);
	std::vector<uint64_t> seen;
	while (sample_pos >= pos_x && closest_v > 0) {
		uint64_t pos = ref_pos;
		if (std::find(seen.begin(), seen.end(), pos) != seen.end()) { if (hang) *hang = true; return vars; }
		seen.push_back(pos);
		closest_v = get_prev_vertex_with_sample(vg, idx, pos, sample_id, ref_pos, sample_pos, ub);
	}
	closest_v = idx->find(ref_pos);                     // start from the ref node at ref_pos (:513)
	SampleInfo sample;
	uint64_t seq_len = 0;
	if (vg->get_sample_from_vertex_if_exists(closest_v, REF, sample)) { seq_len = ref_pos - sample.index; ref_pos = sample.index; }
	else err(log, "reference node is expected to be found!");
	sample_pos = sample_pos - seq_len;
	VariantGraph::PathIterator it = vg->find(closest_v, sample_id);
	std::string cur_ref;
	while (!it.done()) {
		if (sample_pos >= pos_y) break;
		Graph::vertex cur_v = (*it)->vertex_id;
		Variant var;
		uint64_t l = (*it)->length;
		uint64_t next_ref_pos = ref_pos + l;
		uint64_t next_sample_pos = sample_pos + l;
		std::string next_ref;
		VariantGraph::BfsIterator bfs_it = vg->find((*it)->vertex_id, 1);
		++bfs_it;
		while (!bfs_it.done()) {          // last ref-carrying neighbour wins (:541-549)
			Graph::vertex v = (*bfs_it)->vertex_id;
			if (vg->get_sample_from_vertex_if_exists(v, REF, sample)) { next_ref_pos = sample.index; next_ref = vg->get_sequence(*(*bfs_it)); }
			++bfs_it;
		}
		if (sample_pos > pos_x && vg->get_sample_from_vertex_if_exists(cur_v, sample_id, sample)) {
			std::string alt;
			if (ref_pos == next_ref_pos) {                                   // insertion :556-562
				cur_ref = "";
				alt = vg->get_sequence(*(*it));
				var.var_pos = ref_pos; var.var_pos_set = true;
			} else if (vg->get_sample_from_vertex_if_exists(cur_v, REF, sample)) {   // deletion :564-573
				alt = "";
				vg->get_sample_from_vertex_if_exists(cur_v, sample_id, sample);
				var.var_pos = sample.index; var.var_pos_set = true;
				Graph::vertex v = idx->find(ref_pos - 1);
				cur_ref = vg->get_sequence(vg->get_vertex(v));
			} else {                                                         // substitution :574-579
				alt = vg->get_sequence(*(*it));
				vg->get_sample_from_vertex_if_exists(cur_v, sample_id, sample);
				var.var_pos = sample.index; var.var_pos_set = true;
			}
			var.alt = alt;
			var.ref = cur_ref;
			get_samples((*it), vg, var.samples);
			vars.push_back(var);
		}
		cur_ref = next_ref;
		ref_pos = next_ref_pos;
		sample_pos = next_sample_pos;
		++it;
	}
	say(log, "Number of variants get_sample_var_in_sample: " + std::to_string(vars.size()) + "\n");
	if (print) dump_vars(vars, outfile);
	return vars;
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
