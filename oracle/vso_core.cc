// ORACLE — test infrastructure only (see vso.h).  Restatement of the construct side of
// include/variant_graph.h and of include/index.h.  Control flow follows the reference branch by
// branch (citations inline) because neighbour-set iteration order, vertex-id assignment order and
// the dummy-vertex rules are all observable in query output.
#include "vso.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <sys/stat.h>
#include <zlib.h>

namespace vso {

static constexpr uint32_t kVertexesInBlock = 200000;   // variant_graph.h:44
[[noreturn]] static void die(const std::string& m) { throw std::runtime_error("vso: " + m); }

char map_int(uint8_t base) {
	switch (base) { case 0: return 'A'; case 2: return 'T'; case 1: return 'C'; case 3: return 'G'; case 4: return 'N'; default: return (char)5; }
}
uint8_t map_base(char base) {
	switch (base) { case 'A': return 0; case 'T': return 2; case 'C': return 1; case 'G': return 3; case 'N': return 4; default: return 5; }
}

void read_fasta(const std::string& fasta_file, std::string& chr, std::string& ref) {
	std::ifstream stream(fasta_file);
	if (!stream.good()) die("Failed to open input fasta file: " + fasta_file);
	bool found_ref = false;
	std::string line;
	while (getline(stream, line)) {
		if (line.empty()) continue;      // reference: line.at(0) would throw on an empty line
		if (line[0] == '>') {
			if (found_ref) die("Found multiple references in the fasta file");
			std::stringstream line_stream(line);
			getline(line_stream, chr, ' ');
			chr = chr.substr(1);
			found_ref = true;
		} else ref.append(line);
	}
}

// ------------------------------------------------------------------ minimal VCF reader
// Reproduces what add_vcfs consumes from vcflib: CHROM, POS, REF, ALT list, and the name-sorted
// per-sample GT string (vcflib/Variant.h:54 keeps samples in a std::map).
namespace {
struct VcfReader {
	gzFile f = nullptr;
	std::vector<std::string> sample_names;
	std::string line;
	bool open(const std::string& path) {
		f = gzopen(path.c_str(), "rb");
		if (!f) return false;
		gzbuffer(f, 1 << 20);
		while (read_line()) {
			if (line.rfind("##", 0) == 0) continue;
			if (line.rfind("#CHROM", 0) == 0) {
				std::vector<std::string> t = split(line, '\t');
				for (size_t i = 9; i < t.size(); i++) sample_names.push_back(t[i]);
				return true;
			}
			break;
		}
		return false;
	}
	~VcfReader() { if (f) gzclose(f); }
	bool read_line() {
		line.clear();
		char buf[1 << 16];
		while (gzgets(f, buf, sizeof buf)) {
			size_t n = strlen(buf);
			line.append(buf, n);
			if (n && buf[n - 1] == '\n') { line.pop_back(); if (!line.empty() && line.back() == '\r') line.pop_back(); return true; }
		}
		return !line.empty();
	}
	static std::vector<std::string> split(const std::string& s, char d) {
		std::vector<std::string> out; size_t b = 0;
		while (true) { size_t e = s.find(d, b); out.push_back(s.substr(b, e == std::string::npos ? e : e - b)); if (e == std::string::npos) break; b = e + 1; }
		return out;
	}
	bool next(VcfRecord& r) {
		while (read_line()) {
			if (line.empty() || line[0] == '#') continue;
			std::vector<std::string> t = split(line, '\t');
			if (t.size() < 8) continue;
			r.chrom = t[0]; r.pos = atoll(t[1].c_str()); r.ref = t[3]; r.alts = split(t[4], ',');
			r.samples.clear();
			if (t.size() > 9) {
				std::vector<std::string> fmt = split(t[8], ':');
				int gt_i = -1;
				for (size_t i = 0; i < fmt.size(); i++) if (fmt[i] == "GT") gt_i = (int)i;
				for (size_t i = 9; i < t.size() && i - 9 < sample_names.size(); i++) {
					if (t[i] == ".") continue;
					std::vector<std::string> sv = split(t[i], ':');
					std::string gt = (gt_i >= 0 && (size_t)gt_i < sv.size()) ? sv[gt_i] : std::string();
					r.samples.push_back(std::make_pair(sample_names[i - 9], gt));
				}
				std::sort(r.samples.begin(), r.samples.end());
			}
			return true;
		}
		return false;
	}
};

// GT interpretation shared by use_bit_vector_encoding (:584-606) and add_vcfs (:666-705).
bool gt_carries(const std::string& gt_info, bool& phase, bool& gt1, bool& gt2) {
	phase = false; gt1 = false; gt2 = false;
	if (gt_info.size() == 3) {
		int first = gt_info[0] - '0'; char part = gt_info[1]; int second = gt_info[2] - '0';
		if (first > 0 || second > 0) {
			gt1 = first > 0; gt2 = second > 0;
			if (part == '|') phase = true; else if (part == '/') phase = false; else die(std::string("Unknown phase: ") + part);
			return true;
		}
	} else if (gt_info.size() == 1) {
		int present = 0;
		if (gt_info[0] >= '0' && gt_info[0] <= '9') present = gt_info[0] - '0';
		if (present) { gt1 = true; gt2 = false; return true; }
	}
	return false;
}
}  // namespace

// ------------------------------------------------------------------ construct
bool VariantGraph::detect_encoding(const std::string& vcf_file) {   // variant_graph.h:568-617
	VcfReader vr;
	if (!vr.open(vcf_file)) die("can't open vcf " + vcf_file);
	VcfRecord var; uint32_t cnt = 1; float density = 0;
	while (vr.next(var) && cnt < 100) {
		uint32_t n = 0;
		for (const auto& s : var.samples) {
			const std::string& g = s.second;
			if (g.size() == 3) { if (g[0] - '0' > 0 || g[2] - '0' > 0) n++; }
			else if (g.size() == 1) { if (g[0] >= '1' && g[0] <= '9') n++; }
		}
		float cur = n / (float)vr.sample_names.size();
		density = density > cur ? density : cur;
		cnt++;
	}
	return density > 0.05;
}

void VariantGraph::init_ref(const std::string& ref) {            // :331-354
	ref_length = ref.size();
	sampleid_map.insert(std::make_pair("ref", (uint32_t)sampleid_map.size()));
	idsample_map.insert(std::make_pair(0u, "ref"));
	SampleStruct s = {0, 0, 0, 0};
	Vertex* v = add_vertex(ref, 1, 0, s);
	update_idx_vertex_id_map(*v);
}

VariantGraph::VariantGraph(const std::string& ref_file, const std::string& vcf_file, const std::string& pfx,
                           const ConstructOpts& o)
	: prefix(pfx), opts(o), topology(o.cqf_log2_slots, o.use_ref_gqf) {
	std::string ref;
	read_fasta(ref_file, chr, ref);
	use_bit_vector = o.force_encoding < 0 ? detect_encoding(vcf_file) : (o.force_encoding == 1);
	init_ref(ref);
	// add_vcfs :619-733
	VcfReader vr;
	if (!vr.open(vcf_file)) die("can't open vcf " + vcf_file);
	set_sample_names(vr.sample_names);
	VcfRecord var;
	while (vr.next(var)) add_record(var);
	finish_construct();
}

VariantGraph::VariantGraph(const std::string& c, const std::string& ref_seq, const std::string& pfx,
                           const ConstructOpts& o, bool bitvec)
	: chr(c), prefix(pfx), use_bit_vector(bitvec), opts(o), topology(o.cqf_log2_slots, o.use_ref_gqf) {
	init_ref(ref_seq);
}

void VariantGraph::set_sample_names(const std::vector<std::string>& names) {   // :628-632
	for (const auto& sample : names) {
		uint32_t id = (uint32_t)sampleid_map.size();
		if (sampleid_map.insert(std::make_pair(sample, id)).second) idsample_map.insert(std::make_pair(id, sample));
	}
	num_samples += sampleid_map.size();
}

void VariantGraph::add_record(const VcfRecord& var) {                           // :638-729
	num_vars += 1;
	bool chr_ok = (var.chrom == chr) || (var.chrom.size() >= 3 && var.chrom.substr(3) == chr);   // :559-566
	bool seq_ok = false;
	if (chr_ok && var.pos >= 1 && (uint64_t)var.pos <= ref_length) {
		uint64_t st = (uint64_t)var.pos - 1;
		seq_ok = st + var.ref.size() <= seq_buffer.size() && var.ref == get_sequence(st, (uint32_t)var.ref.size());
	}
	if (!seq_ok) return;   // "Unsupported mutation" :641-647
	for (const auto& alt : var.alts) {
		std::vector<SampleStruct> sample_list;
		bool acgt = !alt.empty();
		for (char c : alt) if (c != 'A' && c != 'C' && c != 'T' && c != 'G') { acgt = false; break; }
		if (acgt) {
			for (const auto& sample : var.samples) {
				bool phase, gt1, gt2;
				if (gt_carries(sample.second, phase, gt1, gt2)) {
					auto it = sampleid_map.find(sample.first);
					if (it == sampleid_map.end()) die("Unknown sample: " + sample.first);
					sample_list.push_back(SampleStruct{it->second, phase, gt1, gt2});
				}
			}
		}
		add_allele(var.ref, alt, (uint64_t)var.pos, sample_list);
	}
}

// tail of the per-alt loop of add_vcfs (:723-727); also the entry point of the synthetic
// generators, which build `sample_list` (name-sorted carriers) without VCF text.
void VariantGraph::add_allele(const std::string& ref, const std::string& alt, uint64_t pos, std::vector<SampleStruct>& sample_list) {
	if (!sample_list.empty()) {
		num_mutations += 1;
		num_mutations_samples += sample_list.size();
		add_mutation(ref, alt, pos, sample_list);
	}
}

void VariantGraph::finish_construct() {
	if (opts.fix_sample_indexes) fix_sample_indexes();
}

Vertex* VariantGraph::create_vertex(uint64_t id, uint64_t offset, uint64_t length, uint32_t class_id,
                                    const std::vector<SampleInfo>& samples) {   // :1068-1113
	if (id != vertices.size()) die("vertex id out of sequence");
	vertices.emplace_back();
	Vertex* v = &vertices.back();
	v->vertex_id = (uint32_t)id; v->offset = (uint32_t)offset; v->length = (uint32_t)length;
	if (use_bit_vector) { v->has_class = true; v->class_id = class_id; }
	for (const auto& sample : samples) {
		SampleInfo s; s.index = sample.index;
		if (!use_bit_vector) { s.has_sid = 1; s.sample_id = sample.sample_id; }
		s.phase = sample.phase; s.gt1 = sample.gt1; s.gt2 = sample.gt2;
		v->s_info.push_back(s);
	}
	return v;
}

Vertex* VariantGraph::add_vertex(const std::string& seq, uint64_t index, uint32_t class_id, const SampleStruct& sample) {   // :757-785
	uint64_t start_offset = seq_length;
	seq_buffer.resize(seq_buffer.size() + seq.size());
	for (const auto c : seq) { seq_buffer[seq_length] = map_base(c); seq_length++; }
	SampleInfo s; s.index = (uint32_t)index;
	if (!use_bit_vector) { s.has_sid = 1; s.sample_id = sample.sample_id; }
	s.phase = sample.phase; s.gt1 = sample.gt1; s.gt2 = sample.gt2;
	std::vector<SampleInfo> samples = {s};
	Vertex* v = create_vertex(num_vertices, start_offset, seq.size(), class_id, samples);
	num_vertices++;
	return v;
}

void VariantGraph::add_sample_vector(const BitVec& vector, uint64_t class_id) {   // :787-801
	if (class_id < 1) die("Sample class is smaller than 1.");
	sample_vector.resize(sample_vector.nbits + num_samples);
	uint64_t start_idx = (class_id - 1) * num_samples;
	for (uint32_t i = 0; i < num_samples / 64 * 64; i += 64) sample_vector.set_int(start_idx + i, vector.get_int(i, 64), 64);
	if (num_samples % 64)
		sample_vector.set_int(start_idx + num_samples / 64 * 64, vector.get_int(num_samples / 64 * 64, num_samples % 64), num_samples % 64);
}

uint32_t VariantGraph::find_sample_vector_or_add(const std::vector<SampleStruct>& sample_list) {   // :803-832
	BitVec vector; vector.resize(num_samples);
	for (const auto& sample : sample_list) vector.set(sample.sample_id, 1);
	// MurmurHash64A over vector.capacity()/8 bytes: whole 64-bit words, seed 2038074743 (:811-812)
	uint64_t vec_hash = murmur_hash_64a(vector.w.data(), (int)(vector.w.size() * 8), 2038074743u);
	auto it = sampleclass_map.find(vec_hash);
	if (it == sampleclass_map.end()) {
		uint32_t class_id = (uint32_t)sampleclass_map.size() + 1;
		sampleclass_map.insert(std::make_pair(vec_hash, class_id));
		add_sample_vector(vector, class_id);
		return class_id;
	}
	return it->second;
}

uint32_t VariantGraph::get_popcnt(uint32_t class_id) const {   // :1023-1041
	if (class_id == 0) return 1;
	uint64_t start_idx = (uint64_t)(class_id - 1) * num_samples, popcnt = 0;
	for (uint32_t i = 0; i < num_samples / 64 * 64; i += 64) popcnt += __builtin_popcountll(sample_vector.get_int(start_idx + i, 64));
	if (num_samples % 64) popcnt += __builtin_popcountll(sample_vector.get_int(start_idx + num_samples / 64 * 64, num_samples % 64));
	return (uint32_t)popcnt;
}

uint32_t VariantGraph::get_sample_id(uint32_t class_id, uint32_t index) const {   // :902-942
	if (class_id == 0) return 0;
	uint64_t start_idx = (uint64_t)(class_id - 1) * num_samples;
	uint32_t rank = index + 1;
	for (uint32_t i = 0; i < num_samples / 64 * 64; i += 64) {
		uint64_t word = sample_vector.get_int(start_idx + i, 64);
		uint32_t pc = (uint32_t)__builtin_popcountll(word);
		if (pc >= rank) { for (uint32_t k = 1; k < rank; k++) word &= word - 1; return (uint32_t)__builtin_ctzll(word) + i; }
		rank -= pc;
	}
	if (num_samples % 64) {
		uint64_t word = sample_vector.get_int(start_idx + num_samples / 64 * 64, num_samples % 64);
		uint32_t pc = (uint32_t)__builtin_popcountll(word);
		if (pc >= rank) { for (uint32_t k = 1; k < rank; k++) word &= word - 1; return (uint32_t)__builtin_ctzll(word) + (uint32_t)(num_samples / 64 * 64); }
		die("Index passed is outside the bounds for sample class");
	}
	return UINT32_MAX;
}

uint32_t VariantGraph::get_sample_id(const Vertex& v, uint32_t index) const {   // :875-880
	const SampleInfo& s = v.s_info[index];
	return s.has_sid ? s.sample_id : get_sample_id(v.class_id, index);
}

std::vector<uint32_t> VariantGraph::get_sample_ids(uint32_t class_id) const {   // :944-1006 (both modes list the set bits)
	std::vector<uint32_t> ids;
	if (class_id == 0) { ids.push_back(0); return ids; }
	uint64_t start_idx = (uint64_t)(class_id - 1) * num_samples;
	for (uint64_t i = 0; i < num_samples; i += 64) {
		unsigned len = (unsigned)std::min<uint64_t>(64, num_samples - i);
		uint64_t word = sample_vector.get_int(start_idx + i, len);
		while (word) { ids.push_back((uint32_t)(__builtin_ctzll(word) + i)); word &= word - 1; }
	}
	return ids;
}

std::string VariantGraph::get_sample_phasing(const Vertex& v, uint32_t index) const {   // :882-900
	const SampleInfo& s = v.s_info[index];
	std::string p;
	p += s.gt1 ? "1" : "0"; p += s.phase ? "|" : "/"; p += s.gt2 ? "1" : "0";
	return p;
}

std::string VariantGraph::get_sample_name(uint32_t id) const {   // :1230-1236
	auto it = idsample_map.find(id);
	if (it == idsample_map.end()) die("Unknown sample id: " + std::to_string(id));
	return it->second;
}

uint32_t VariantGraph::sample_id_of(const std::string& name) const {
	auto it = sampleid_map.find(name);
	if (it == sampleid_map.end()) die("Sample not found: " + name);
	return it->second;
}

std::string VariantGraph::get_sequence(const Vertex& v) const {
	std::string seq; seq.reserve(v.length);
	for (uint64_t i = v.offset; i < (uint64_t)v.offset + v.length; i++) seq += map_int(seq_buffer[i]);
	return seq;
}
std::string VariantGraph::get_sequence(uint64_t start, uint32_t length) const {
	std::string seq; seq.reserve(length);
	for (uint64_t i = start; i < start + length; i++) seq += map_int(seq_buffer[i]);
	return seq;
}

void VariantGraph::update_idx_vertex_id_map(const Vertex& v) {   // :1280-1287
	for (size_t i = 0; i < v.s_info.size(); i++)
		if (get_sample_id(v, (uint32_t)i) == 0) idx_vertex_id[v.s_info[i].index] = v.vertex_id;
}

bool VariantGraph::get_sample_from_vertex_if_exists(Graph::vertex v, uint32_t sample_id, SampleInfo& sample) const {   // :1296-1326
	const Vertex& cur_vertex = get_vertex(v);
	if (is_bit_vector(cur_vertex)) {
		uint32_t idx = 0;
		auto sample_ids = get_sample_ids(cur_vertex.class_id);
		for (auto id : sample_ids) {
			if (id == sample_id) {
				if (idx >= cur_vertex.s_info.size()) die("s_info / class popcount mismatch at vertex " + std::to_string(v));
				sample = cur_vertex.s_info[idx];
				return true;
			}
			idx++;
		}
	} else {
		for (size_t i = 0; i < cur_vertex.s_info.size(); i++)
			if (get_sample_id(cur_vertex, (uint32_t)i) == sample_id) { sample = cur_vertex.s_info[i]; return true; }
	}
	return false;
}

bool VariantGraph::get_sample_from_vertex_if_exists(Graph::vertex v, const std::string& sample_id, SampleInfo& sample) const {
	return get_sample_from_vertex_if_exists(v, sample_id_of(sample_id), sample);
}

bool VariantGraph::get_neighbor_vertex(Graph::vertex id, uint32_t sample_id, Graph::vertex* v) const {   // :1402-1451
	uint32_t min_idx = UINT32_MAX;
	for (const auto v_id : topology.out_neighbors(id)) {
		const Vertex& vertex = get_vertex(v_id);
		if (is_bit_vector(vertex)) {
			uint32_t idx = 0;
			auto sample_ids = get_sample_ids(vertex.class_id);
			for (auto s_id : sample_ids) {
				if (s_id != 0 && s_id == sample_id) { *v = v_id; return true; }
				else if (s_id == 0) {
					const SampleInfo& s = vertex.s_info[idx];
					if (min_idx > s.index) { *v = v_id; min_idx = s.index; }
				}
				idx++;
			}
		} else {
			for (size_t i = 0; i < vertex.s_info.size(); i++) {
				const SampleInfo& s = vertex.s_info[i];
				uint32_t s_id = get_sample_id(vertex, (uint32_t)i);
				if (s_id != 0 && s_id == sample_id) { *v = v_id; return true; }
				else if (s_id == 0) { if (min_idx > s.index) { *v = v_id; min_idx = s.index; } }
			}
		}
	}
	return *v != 0;
}

void VariantGraph::add_sample_to_vertex(Graph::vertex id, uint64_t sample_idx, const SampleStruct& sample) {   // :1453-1464
	SampleInfo s; s.index = (uint32_t)sample_idx;
	if (!use_bit_vector) { s.has_sid = 1; s.sample_id = sample.sample_id; }
	s.phase = sample.phase; s.gt1 = sample.gt1; s.gt2 = sample.gt2;
	get_mutable_vertex(id).s_info.push_back(s);
}

void VariantGraph::validate_ref_path_edge(Graph::vertex src, Graph::vertex dest) const {   // :1483-1494
	SampleInfo a, b;
	if (!get_sample_from_vertex_if_exists(src, 0u, a) && !get_sample_from_vertex_if_exists(dest, 0u, b)) {
		if (a.index >= b.index) die("Source ref index is not smaller than dest ref index");
	}
}

bool VariantGraph::update_vertex_sample_class(Graph::vertex vertex_id, const std::vector<SampleStruct>& sample_list) {   // :834-873
	std::map<uint32_t, SampleStruct> sample_indexes;
	std::vector<SampleStruct> list;
	for (const auto& sample : sample_list) sample_indexes.insert(std::make_pair(sample.sample_id, sample));
	uint64_t ref_index = 0;
	{
		const Vertex& v = get_vertex(vertex_id);
		for (int i = (int)v.s_info.size() - 1; i >= 0; i--) {
			SampleStruct s = {get_sample_id(v, (uint32_t)i), 0, 0, 0};   // carriers already on the vertex lose their GT (:845)
			if (s.sample_id == 0) ref_index = v.s_info[i].index;
			sample_indexes.insert(std::make_pair(s.sample_id, s));
		}
	}
	for (const auto& sample : sample_indexes) list.emplace_back(sample.second);
	uint32_t class_id = find_sample_vector_or_add(list);
	Vertex& v = get_mutable_vertex(vertex_id);
	v.class_id = class_id;
	if (list.size() != get_popcnt(class_id)) return false;
	v.s_info.clear();
	for (const auto& sample : list) {
		if (sample.sample_id == 0) add_sample_to_vertex(vertex_id, ref_index, sample);
		else add_sample_to_vertex(vertex_id, 0, sample);
	}
	return true;
}

void VariantGraph::split_vertex(uint64_t vertex_id, uint64_t pos, Graph::vertex* new_vertex) {   // :1115-1160
	const Vertex& cv = get_vertex((Graph::vertex)vertex_id);
	uint32_t cur_offset = cv.offset, cur_length = cv.length;
	SampleInfo cur_s0 = cv.s_info[0];
	if (pos > cur_length) die("Split position is greater than vertex length.");
	uint64_t offset = cur_offset + pos - 1;
	uint64_t length = cur_length - pos + 1;
	SampleInfo s = cur_s0; s.index = (uint32_t)(cur_s0.index + pos - 1);
	std::vector<SampleInfo> samples = {s};
	Vertex* v = create_vertex(num_vertices, offset, length, 0, samples);
	*new_vertex = v->vertex_id;
	if (cur_s0.index != v->s_info[0].index) update_idx_vertex_id_map(*v);
	get_mutable_vertex((Graph::vertex)vertex_id).length = (uint32_t)(cur_length - length);
	for (const auto n : topology.out_neighbors((Graph::vertex)vertex_id)) {
		validate_ref_path_edge(*new_vertex, n);
		topology.add_edge(*new_vertex, n);
		topology.remove_edge((Graph::vertex)vertex_id, n);
	}
	validate_ref_path_edge((Graph::vertex)vertex_id, *new_vertex);
	topology.add_edge((Graph::vertex)vertex_id, *new_vertex);
	num_vertices++;
}

void VariantGraph::split_vertex(uint64_t vertex_id, uint64_t pos1, uint64_t pos2, Graph::vertex* n1, Graph::vertex* n2) {   // :1162-1167
	split_vertex(vertex_id, pos1, n1);
	split_vertex(*n1, pos2 - pos1 + 1, n2);
}

void VariantGraph::add_mutation(std::string ref, std::string alt, uint64_t pos, std::vector<SampleStruct>& sample_list) {   // :1509-1881
	MUT mutation;
	if (ref.size() == alt.size()) mutation = SUBSTITUTION;
	else if (ref.size() > alt.size()) mutation = DELETION;
	else mutation = INSERTION;
	if (mutation == INSERTION) { pos = pos + ref.size(); alt = alt.substr(ref.size()); }
	else if (mutation == DELETION) { pos = pos + alt.size(); ref = ref.substr(alt.size()); }

	// vertex of the backbone holding @pos (:1535-1546)
	auto ref_idx_itr = idx_vertex_id.lower_bound(pos);
	if (ref_idx_itr == idx_vertex_id.end() || ref_idx_itr->first != pos) {
		--ref_idx_itr;
		if (ref_idx_itr->first >= pos) die("The prev ref vertex has an index greater than pos.");
	}
	const uint64_t ref_vertex_idx = ref_idx_itr->first;
	Graph::vertex ref_vertex_id = (Graph::vertex)ref_idx_itr->second;
	const uint64_t rv_length = get_vertex(ref_vertex_id).length;   // the reference copies the vertex here: pre-split values
	const uint64_t rv_offset = get_vertex(ref_vertex_id).offset;

	// helper for the three "mutation spans one or more vertexes" branches (:1603-1618, :1648-1663, :1830-1844)
	auto span_forward = [&](Graph::vertex* next_ref_vertex_id, bool has_else) {
		auto temp_itr = idx_vertex_id.lower_bound(ref_vertex_idx);
		uint64_t nlen; Graph::vertex nid;
		do {
			++temp_itr;
			if (temp_itr == idx_vertex_id.end()) die("mutation runs past the end of the reference");
			nid = (Graph::vertex)temp_itr->second; nlen = get_vertex(nid).length;
		} while (temp_itr->first + nlen < pos + ref.size());
		if (temp_itr->first + nlen == pos + ref.size()) get_neighbor_vertex((Graph::vertex)temp_itr->second, 0, next_ref_vertex_id);
		else if (!has_else || temp_itr->first + nlen < pos + ref.size())   // the `<` arm is unreachable after the loop
			split_vertex(temp_itr->second, pos + ref.size() - temp_itr->first + 1, next_ref_vertex_id);
		else *next_ref_vertex_id = nid;
	};

	if (mutation == SUBSTITUTION) {
		Graph::vertex prev_ref_vertex_id = 0, next_ref_vertex_id = 0;
		if (ref_vertex_idx == pos && rv_length == ref.size()) {                       // :1552-1570
			split_vertex(ref_vertex_id, 1, &next_ref_vertex_id);                        // dummy vertex
			prev_ref_vertex_id = ref_vertex_id; ref_vertex_id = next_ref_vertex_id;
			get_neighbor_vertex(ref_vertex_id, 0, &next_ref_vertex_id);
		} else if (ref_vertex_idx == pos && rv_length > ref.size()) {                 // :1571-1590
			split_vertex(ref_vertex_id, 1, &next_ref_vertex_id);
			prev_ref_vertex_id = ref_vertex_id; ref_vertex_id = next_ref_vertex_id;
			split_vertex(ref_vertex_id, ref.size() + 1, &next_ref_vertex_id);
		} else if (ref_vertex_idx == pos && rv_length < ref.size()) {                 // :1591-1618
			if (rv_length > 1) split_vertex(ref_vertex_id, 1, &next_ref_vertex_id);
			prev_ref_vertex_id = ref_vertex_id; ref_vertex_id = next_ref_vertex_id;
			span_forward(&next_ref_vertex_id, true);
		} else if (ref_vertex_idx < pos && ref_vertex_idx + rv_length > pos + ref.size()) {   // :1628-1636
			uint64_t split_pos = pos - rv_offset;
			split_vertex(ref_vertex_id, split_pos, split_pos + ref.size(), &prev_ref_vertex_id, &next_ref_vertex_id);
			std::swap(ref_vertex_id, prev_ref_vertex_id);
		} else if (ref_vertex_idx < pos && ref_vertex_idx + rv_length < pos + ref.size()) {   // :1637-1663
			prev_ref_vertex_id = ref_vertex_id;
			if (rv_length > 1) split_vertex(prev_ref_vertex_id, pos - ref_vertex_idx + 1, &ref_vertex_id);
			span_forward(&next_ref_vertex_id, true);
		} else if (ref_vertex_idx < pos && ref_vertex_idx + rv_length == pos + ref.size()) {  // :1664-1672
			prev_ref_vertex_id = ref_vertex_id;
			split_vertex(prev_ref_vertex_id, pos - ref_vertex_idx + 1, &ref_vertex_id);
			get_neighbor_vertex(ref_vertex_id, 0, &next_ref_vertex_id);
		}
		uint32_t class_id = 0;
		if (use_bit_vector) class_id = find_sample_vector_or_add(sample_list);
		Graph::vertex sv = add_vertex(alt, 0, class_id, sample_list[0])->vertex_id;
		topology.add_edge(prev_ref_vertex_id, sv);
		topology.add_edge(sv, next_ref_vertex_id);
		sample_list.erase(sample_list.begin());
		for (const auto& sample : sample_list) add_sample_to_vertex(sv, 0, sample);
		if (use_bit_vector && (uint32_t)get_vertex(sv).s_info.size() != get_popcnt(get_vertex(sv).class_id))
			die("Num of samples is not equal to num of 1s in the sample class.");
	} else if (mutation == INSERTION) {
		Graph::vertex prev_ref_vertex_id = 0, next_ref_vertex_id = 0;
		if (ref_vertex_idx == pos) {                                                  // :1706-1717
			auto temp_itr = idx_vertex_id.lower_bound(ref_vertex_idx);
			if (temp_itr->first != ref_vertex_idx) die("Vertex id not found in the map");
			if (temp_itr == idx_vertex_id.begin()) die("insertion before the first base");
			--temp_itr;
			prev_ref_vertex_id = (Graph::vertex)temp_itr->second;
			next_ref_vertex_id = ref_vertex_id;
		} else if (ref_vertex_idx < pos && ref_vertex_idx + rv_length > pos) {        // :1718-1723
			split_vertex(ref_vertex_id, pos - ref_vertex_idx + 1, &next_ref_vertex_id);
			prev_ref_vertex_id = ref_vertex_id;
		} else if (ref_vertex_idx + rv_length == pos) {                               // :1724-1728
			prev_ref_vertex_id = ref_vertex_id;
			get_neighbor_vertex(ref_vertex_id, 0, &next_ref_vertex_id);
		} else {                                                                      // :1729-1731
			prev_ref_vertex_id = ref_vertex_id;
		}
		uint32_t class_id = 0;
		if (use_bit_vector) class_id = find_sample_vector_or_add(sample_list);
		Graph::vertex sv = add_vertex(alt, 0, class_id, sample_list[0])->vertex_id;
		topology.add_edge(prev_ref_vertex_id, sv);
		if (next_ref_vertex_id != 0) topology.add_edge(sv, next_ref_vertex_id);
		sample_list.erase(sample_list.begin());
		for (const auto& sample : sample_list) add_sample_to_vertex(sv, 0, sample);
		if (use_bit_vector && (uint32_t)get_vertex(sv).s_info.size() != get_popcnt(get_vertex(sv).class_id))
			die("Num of samples is not equal to num of 1s in the sample class.");
	} else {   // DELETION
		Graph::vertex prev_ref_vertex_id = 0, next_ref_vertex_id = 0;
		auto prev_of_ref_vertex = [&]() {
			auto temp_itr = idx_vertex_id.lower_bound(ref_vertex_idx);
			if (temp_itr->first != ref_vertex_idx) die("Vertex id not found in the map");
			if (temp_itr == idx_vertex_id.begin()) die("deletion before the first base");
			--temp_itr;
			return (Graph::vertex)temp_itr->second;
		};
		if (ref_vertex_idx == pos && rv_length == ref.size()) {                       // :1764-1776
			prev_ref_vertex_id = prev_of_ref_vertex();
			get_neighbor_vertex(ref_vertex_id, 0, &next_ref_vertex_id);
		} else if (ref_vertex_idx == pos && rv_length > ref.size()) {                 // :1777-1789
			prev_ref_vertex_id = prev_of_ref_vertex();
			split_vertex(ref_vertex_id, ref.size() + 1, &next_ref_vertex_id);
		} else if (ref_vertex_idx == pos && rv_length < ref.size()) {                 // :1790-1813
			span_forward(&next_ref_vertex_id, false);
			prev_ref_vertex_id = prev_of_ref_vertex();
		} else if (ref_vertex_idx < pos && ref_vertex_idx + rv_length > pos + ref.size()) {   // :1814-1820
			uint64_t split_pos = pos - rv_offset;
			split_vertex(ref_vertex_id, split_pos, split_pos + ref.size(), &prev_ref_vertex_id, &next_ref_vertex_id);
			std::swap(ref_vertex_id, prev_ref_vertex_id);
		} else if (ref_vertex_idx < pos && ref_vertex_idx + rv_length < pos + ref.size()) {   // :1821-1844
			prev_ref_vertex_id = ref_vertex_id;
			if (rv_length > 1) split_vertex(prev_ref_vertex_id, pos - ref_vertex_idx + 1, &ref_vertex_id);
			span_forward(&next_ref_vertex_id, true);
		} else if (ref_vertex_idx < pos && ref_vertex_idx + rv_length == pos + ref.size()) {  // :1845-1853
			prev_ref_vertex_id = ref_vertex_id;
			split_vertex(prev_ref_vertex_id, pos - ref_vertex_idx + 1, &ref_vertex_id);
			get_neighbor_vertex(ref_vertex_id, 0, &next_ref_vertex_id);
		}
		if (use_bit_vector) {                                                         // :1856-1871
			update_vertex_sample_class(next_ref_vertex_id, sample_list);
			if ((uint32_t)get_vertex(next_ref_vertex_id).s_info.size() != get_popcnt(get_vertex(next_ref_vertex_id).class_id))
				die("Num of samples is not equal to num of 1s in the sample class.");
		} else {
			for (const auto& sample : sample_list) add_sample_to_vertex(next_ref_vertex_id, 0, sample);
		}
		validate_ref_path_edge(prev_ref_vertex_id, next_ref_vertex_id);
		topology.add_edge(prev_ref_vertex_id, next_ref_vertex_id);
	}
}

void VariantGraph::fix_sample_indexes() {   // :1905-1996 ("optimized solution")
	std::unordered_map<uint32_t, int32_t> sampleid_delta;
	for (const auto& sample : idsample_map) sampleid_delta.insert(std::make_pair(sample.first, 0));
	Graph::GraphIterator it(&topology, 0, UINT64_MAX);
	while (!it.done()) {
		Graph::vertex cur_id = *it;
		SampleInfo ref_sample;
		if (get_sample_from_vertex_if_exists(cur_id, 0u, ref_sample)) {
			uint32_t ref_index = ref_sample.index;
			uint32_t cur_len = get_vertex(cur_id).length;
			for (auto neighbor_id : topology.out_neighbors(cur_id)) {
				Vertex& nb = get_mutable_vertex(neighbor_id);
				for (size_t i = 0; i < nb.s_info.size(); ++i) {
					SampleInfo& s = nb.s_info[i];
					uint32_t s_id = get_sample_id(nb, (uint32_t)i);
					if (s_id != 0 && s.index == 0) {
						auto map_it = sampleid_delta.find(s_id);
						if (map_it == sampleid_delta.end()) die("Unknown sample id");
						int32_t delta = map_it->second;
						int32_t sample_index = ref_index + cur_len + delta;
						if (sample_index < 0) die("Sample index is less than 0");
						s.index = sample_index;
					} else if (s_id != 0 && s.index != 0) {
						sampleid_delta[s_id] = s.index - (ref_index + cur_len);
					}
				}
			}
		} else {
			auto nbrs = topology.out_neighbors(cur_id);
			if (nbrs.size() > 1) die("Sample vertex has more than 1 neighbor: " + std::to_string(cur_id));
			for (auto neighbor_id : nbrs) {
				SampleInfo rs;
				if (!get_sample_from_vertex_if_exists(neighbor_id, 0u, rs)) die("Ref vertex not found as a neighbor from sample vertex.");
				const Vertex& cv = get_vertex(cur_id);
				uint32_t cur_length = cv.length;
				for (size_t i = 0; i < cv.s_info.size(); ++i) {
					uint32_t cur_index = cv.s_info[i].index;
					uint32_t s_id = get_sample_id(cv, (uint32_t)i);
					sampleid_delta[s_id] = cur_index + cur_length - rs.index;
				}
			}
		}
		++it;
	}
}

// ------------------------------------------------------------------ (de)serialisation  :366-446, :501-557
void VariantGraph::serialize() {
	if (read_only) die("Serialization not allowed. VariantStore is loaded in READ ONLY mode");
	mkdir(prefix.c_str(), 0755);
	for (uint64_t b = 0; b * kVertexesInBlock < vertices.size(); b++) {
		size_t lo = b * kVertexesInBlock, hi = std::min<size_t>(vertices.size(), lo + kVertexesInBlock);
		if (!codec::write_vertex_block(prefix + "/vertex_list_" + std::to_string(b) + ".proto", &vertices[lo], hi - lo, opts.gzip_level))
			die("Failed to write vertex list.");
	}
	{
		uint64_t mx = 0; std::vector<uint64_t> vals(seq_buffer.begin(), seq_buffer.end());
		for (auto v : vals) mx = std::max(mx, v);
		uint8_t width = std::min<uint8_t>(3, codec::bits_needed(mx));   // seq_buffer starts 3 bits wide (:335); bit_compress only shrinks
		if (!codec::write_int_vector0(prefix + "/seq_buffer.sdsl", vals, width)) die("Failed to serialize seq buffer");
	}
	topology.serialize(prefix);
	if (!codec::write_rrr127(prefix + "/sample_vector.sdsl", sample_vector)) die("Failed to serialize compressed sample vector");
	std::ofstream f(prefix + "/sampleid_map.lst");
	if (!f.good()) die("Failed to open sampleid file");
	f << chr << " " << std::to_string(ref_length) << "\n";
	f << chr << " " << std::to_string(num_samples) << "\n";
	for (const auto& sample : sampleid_map) f << sample.first << " " << sample.second << "\n";
}

VariantGraph::VariantGraph(const std::string& pfx, bool use_ref_gqf) : prefix(pfx), read_only(true), topology(pfx, use_ref_gqf) {
	// every vertex_list_<k>.proto, ordered by k (:371-397)
	for (uint64_t b = 0;; b++) {
		std::string name = prefix + "/vertex_list_" + std::to_string(b) + ".proto";
		struct stat st;
		if (stat(name.c_str(), &st) != 0) break;
		if (!codec::read_vertex_block(name, vertices)) die("Failed to parse vertex list " + name);
	}
	std::vector<uint64_t> sb;
	if (!codec::read_int_vector0(prefix + "/seq_buffer.sdsl", sb)) die("Failed to load seq buffer");
	seq_buffer.assign(sb.begin(), sb.end());
	num_vertices = topology.get_num_vertices() + 1;
	seq_length = seq_buffer.size();
	if (!codec::read_rrr127(prefix + "/sample_vector.sdsl", sample_vector)) die("Failed to load sample vector");
	std::ifstream f(prefix + "/sampleid_map.lst");
	if (!f.good()) die("Failed to open sampleid map file");
	std::string sample; uint32_t id;
	f >> chr >> ref_length;
	f >> sample >> num_samples;
	if (num_samples <= 0) die("Num samples is less or equal to 0.");
	while (f >> sample >> id) { sampleid_map.insert(std::make_pair(sample, id)); idsample_map.insert(std::make_pair(id, sample)); }
	if (num_samples != sampleid_map.size()) die("Num samples is not equal to num entries in samples file.");
	use_bit_vector = !vertices.empty() && !vertices[0].s_info.empty() && !vertices[0].s_info[0].has_sid;
}

// ------------------------------------------------------------------ path iterator  :1999-2036
VariantGraph::PathIterator::PathIterator(const VariantGraph* g, Graph::vertex v, const std::string& sample) {
	vg = g; cur = &vg->get_vertex(v); s_id = vg->sample_id_of(sample); is_done = false;
}
void VariantGraph::PathIterator::operator++() {
	Graph::vertex next_vertex = 0;
	if (!vg->get_neighbor_vertex(cur->vertex_id, s_id, &next_vertex) && next_vertex == 0) is_done = true;
	cur = &vg->get_vertex(next_vertex);
}

// ------------------------------------------------------------------ Index  (index.h)
Index::Index(const VariantGraph* vg) {   // :53-106
	size_bits = vg->get_ref_length();
	std::vector<uint8_t> b(size_bits, 0);
	auto it = vg->find("ref");
	while (!it.done()) {
		uint64_t node_id = (*it)->vertex_id;
		SampleInfo sample;
		if (!vg->get_sample_from_vertex_if_exists((Graph::vertex)node_id, "ref", sample)) die("Ref sample not found in the vertex");
		uint64_t idx = sample.index;
		if (idx < 1 || idx - 1 >= size_bits) die("ref index out of range while building the index");
		if (b[idx - 1] != 1) node_list.push_back((uint32_t)node_id);
		b[idx - 1] = 1;
		++it;
	}
	for (uint64_t i = 0; i < size_bits; i++) if (b[i]) ones.push_back(i);
}

Index::Index(const std::string& prefix) {   // :108-117
	BitVec bv;
	if (!codec::read_rrr127(prefix + "/index.sdsl", bv)) die("can't read index.sdsl");
	size_bits = bv.nbits;
	for (size_t wi = 0; wi < bv.w.size(); wi++) { uint64_t w = bv.w[wi]; while (w) { ones.push_back(wi * 64 + __builtin_ctzll(w)); w &= w - 1; } }
	std::vector<uint64_t> nl;
	if (!codec::read_int_vector0(prefix + "/ref_node_id.sdsl", nl)) die("can't read ref_node_id.sdsl");
	node_list.assign(nl.begin(), nl.end());
}

void Index::serialize(const std::string& prefix) const {   // :174-179
	BitVec bv; bv.resize(size_bits);
	for (auto p : ones) bv.set(p, 1);
	codec::write_rrr127(prefix + "/index.sdsl", bv);
	std::vector<uint64_t> nl(node_list.begin(), node_list.end());
	uint64_t mx = 0; for (auto v : nl) mx = std::max(mx, v);
	codec::write_int_vector0(prefix + "/ref_node_id.sdsl", nl, std::min<uint8_t>(32, codec::bits_needed(mx)));
}

uint64_t Index::rank(uint64_t pos) const { return std::lower_bound(ones.begin(), ones.end(), pos) - ones.begin(); }

Graph::vertex Index::find(uint64_t pos) const {   // :119-133
	if (pos < 1) die("Can't find node corresponding to pos 0");
	if (pos >= size_bits) return node_list[node_list.size() - 1];
	uint64_t node_idx = rank(pos);
	if (node_idx == 0) return node_list[0];
	return node_list[node_idx - 1];
}

Graph::vertex Index::find(uint64_t pos, uint64_t& ref_node_rank) const {   // :135-148
	if (pos >= size_bits) { ref_node_rank = node_list.size() - 1; return node_list[node_list.size() - 1]; }
	uint64_t node_idx = rank(pos);
	if (node_idx == 0) return node_list[0];
	ref_node_rank = node_idx - 1;
	return node_list[node_idx - 1];
}

bool Index::is_empty(uint64_t pos_x, uint64_t pos_y) const {   // :150-166
	if (pos_x < 1) die("Can't find node corresponding to pos 0");
	if (pos_x > size_bits) return true;
	uint64_t r = rank(pos_x);
	// select(r), select(r+1): 1-based.  select past the last one is undefined in sdsl; the oracle
	// (and the engine) define it as "empty".
	if (r < 1 || r + 1 > ones.size()) return true;
	uint64_t index_x = ones[r - 1], index_y = ones[r];
	if (index_x <= pos_x && index_y <= pos_y) return false;
	return true;
}

Graph::vertex Index::previous(uint64_t ref_node_rank) const {   // :168-172
	if (ref_node_rank == 0) return node_list[0];
	return node_list[ref_node_rank - 1];
}

}  // namespace vso
