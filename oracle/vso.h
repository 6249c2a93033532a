// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under variantstore_b200/ may include, link or
// execute this code; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs use it, and only as the checker / reported baseline.
//
// CPU restatement (a "port": the reference itself cannot be built here — no sdsl-lite, no
// protobuf C++, no htslib) of the VariantStore construct + query path:
//   include/graph.h           topology store (CQF key -> in-place neighbour | aux-set pointer)
//   include/variant_graph.h   variation graph build (add_mutation / split_vertex / classes)
//   include/index.h           position index (bit-vector of vertex starts + node_list)
//   include/query.h           operators t4 / t6 / t7 (+ helpers)
// All file:line citations are relative to /root/reference.
//
// Parity status: pinned against the README goldens of the reference (README.md:58-60,93-95);
// the sdsl rrr_vector on-disk layout is restated from memory of sdsl-lite and is UNPINNED.
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

namespace vso {

// ---------------------------------------------------------------- plain bit / int containers
// Stand-ins for sdsl::bit_vector / sdsl::int_vector<> (only the operations the reference uses).
struct BitVec {
	uint64_t nbits = 0;
	std::vector<uint64_t> w;
	void resize(uint64_t n) { nbits = n; w.resize((n + 63) / 64, 0); }
	bool get(uint64_t i) const { return (w[i >> 6] >> (i & 63)) & 1; }
	void set(uint64_t i, bool b) {
		if (b) w[i >> 6] |= (1ULL << (i & 63)); else w[i >> 6] &= ~(1ULL << (i & 63));
	}
	uint64_t get_int(uint64_t pos, unsigned len) const;           // len in 1..64
	void set_int(uint64_t pos, uint64_t v, unsigned len);
};

// ---------------------------------------------------------------- vertex records
// variantgraphvertex.proto:6-26
struct SampleInfo {
	uint32_t index = 0;
	uint32_t sample_id = 0;   // valid iff has_sid (explicit-id encoding)
	uint8_t has_sid = 0;
	uint8_t phase = 0, gt1 = 0, gt2 = 0;
};
struct Vertex {
	uint32_t vertex_id = 0, offset = 0, length = 0;
	uint32_t class_id = 0;    // valid iff has_class (bit-vector encoding)
	bool has_class = false;
	std::vector<SampleInfo> s_info;
};
struct SampleStruct { uint32_t sample_id; bool phase, gt1, gt2; };  // variant_graph.h:77-82

// ---------------------------------------------------------------- CQF-semantics adjacency store
// What Graph needs from CQF<KeyObject> (gqf_cpp.h:34-112): key -> (1 value bit, count).
class AdjStore {
public:
	virtual ~AdjStore() {}
	virtual uint64_t query(uint64_t key, uint64_t* value_bit) const = 0;   // qf_query
	virtual int insert(uint64_t key, uint64_t value, uint64_t count) = 0;  // qf_insert
	virtual int remove(uint64_t key, uint64_t value) = 0;                  // qf_delete_key_value
	virtual uint64_t ndistinct() const = 0;                                // metadata->ndistinct_elts
	virtual bool serialize(const std::string& path) const = 0;             // qf_serialize
	// enumerate (key, value bit, count) in CQF iteration order (ascending hash)
	virtual void enumerate(std::vector<std::array<uint64_t, 3>>& out) const = 0;
};
std::unique_ptr<AdjStore> make_port_adjstore(unsigned log2_slots);
std::unique_ptr<AdjStore> load_port_adjstore(const std::string& path);   // parses adj_list.cqf
// Real reference gqf compiled from /root/reference/src/gqf (oracle/_ref/libgqf_ref.so); nullptr
// when that library is not built.
std::unique_ptr<AdjStore> make_ref_adjstore(unsigned log2_slots);
std::unique_ptr<AdjStore> load_ref_adjstore(const std::string& path);

// ---------------------------------------------------------------- Graph  (graph.h:40-137)
class Graph {
public:
	typedef uint32_t vertex;
	typedef std::unordered_set<vertex> vertex_set;

	explicit Graph(unsigned log2_slots = 25, bool use_ref_gqf = false);
	Graph(const std::string& prefix, bool use_ref_gqf);      // graph.h:149-172
	int add_edge(vertex s, vertex d);                        // graph.h:210-240
	int remove_edge(vertex s, vertex d);                     // graph.h:242-263
	vertex_set out_neighbors(vertex v) const;                // graph.h:265-280 (by value)
	uint32_t out_degree(vertex v) const;                     // graph.h:282-296
	bool is_edge(vertex s, vertex d) const;                  // graph.h:298-316
	uint32_t get_num_vertices() const { return (uint32_t)adj->ndistinct(); }
	uint32_t get_num_edges() const { return num_edges; }
	void serialize(const std::string& prefix) const;         // graph.h:179-208

	// graph.h:79-93, :394-459  BFS with radius
	class GraphIterator {
	public:
		GraphIterator(const Graph* g, vertex v, uint64_t radius);
		vertex operator*() const { return cur; }
		void operator++();
		bool done() const { return is_done; }
	private:
		vertex cur; uint64_t r; bool is_done; const Graph* g;
		std::vector<std::pair<vertex, uint64_t>> q; size_t qh = 0;   // std::queue stand-in
		std::unordered_set<vertex> visited;
	};

	std::unique_ptr<AdjStore> adj;
	std::vector<vertex_set> aux_vertex_list;
	uint32_t num_edges = 0;
};

// ---------------------------------------------------------------- VariantGraph
struct ConstructOpts {
	unsigned cqf_log2_slots = 25;      // graph.h:29 DEFAULT_SIZE (1<<25); tests use a smaller table
	bool use_ref_gqf = false;          // drive the real gqf library instead of the port
	bool fix_sample_indexes = true;    // variant_graph.h:1883-1997 (not on the t4/t6/t7 path)
	int force_encoding = -1;           // -1 auto (variant_graph.h:568-617), 0 explicit ids, 1 classes
	int gzip_level = -1;               // protobuf GzipOutputStream default
};

struct VcfRecord {                   // what vcflib::Variant exposes to add_vcfs
	std::string chrom; int64_t pos = 0; std::string ref;
	std::vector<std::string> alts;
	// name-sorted (vcflib/Variant.h:54 std::map) list of (sample name, GT string)
	std::vector<std::pair<std::string, std::string>> samples;
};

class VariantGraph {
public:
	// construct (variant_graph.h:323-364)
	VariantGraph(const std::string& ref_file, const std::string& vcf_file, const std::string& prefix,
	             const ConstructOpts& o);
	// programmatic construct used by the synthetic generators (no VCF text): the caller feeds
	// records through begin_records / add_record / end_records.
	VariantGraph(const std::string& chr, const std::string& ref_seq, const std::string& prefix,
	             const ConstructOpts& o, bool use_bit_vector);
	// load from disk (variant_graph.h:366-446)
	VariantGraph(const std::string& prefix, bool use_ref_gqf = false);

	void set_sample_names(const std::vector<std::string>& names);   // variant_graph.h:628-632
	void add_record(const VcfRecord& r);                            // body of add_vcfs loop :638-729
	void add_allele(const std::string& ref, const std::string& alt, uint64_t pos, std::vector<SampleStruct>& carriers);  // :723-727
	void count_record() { num_vars += 1; }
	void finish_construct();                                        // :359-363
	void serialize();                                               // :501-557

	uint64_t get_num_vertices() const { return topology.get_num_vertices(); }
	uint64_t get_num_edges() const { return topology.get_num_edges(); }
	uint64_t get_seq_length() const { return seq_length; }
	uint64_t get_ref_length() const { return ref_length; }
	uint64_t get_num_sample_classes() const { return sampleclass_map.size(); }
	const std::string& get_chr() const { return chr; }
	std::string get_sample_name(uint32_t id) const;
	bool has_sample(const std::string& name) const { return sampleid_map.count(name) != 0; }

	const Vertex& get_vertex(Graph::vertex id) const { return vertices[id]; }
	Vertex& get_mutable_vertex(Graph::vertex id) { return vertices[id]; }
	std::string get_sequence(const Vertex& v) const;                        // :1261-1268
	std::string get_sequence(uint64_t start, uint32_t length) const;        // :1270-1277
	bool get_sample_from_vertex_if_exists(Graph::vertex v, uint32_t sample_id, SampleInfo& s) const;   // :1296-1326
	bool get_sample_from_vertex_if_exists(Graph::vertex v, const std::string& sample, SampleInfo& s) const;  // :1328-1339
	uint32_t get_sample_id(const Vertex& v, uint32_t index) const;          // :875-880
	std::string get_sample_phasing(const Vertex& v, uint32_t index) const;  // :882-900
	bool get_neighbor_vertex(Graph::vertex id, uint32_t sample_id, Graph::vertex* v) const;   // :1402-1451
	std::vector<uint32_t> get_sample_ids(uint32_t sampleclass_id) const;    // :944-1006
	uint32_t sample_id_of(const std::string& name) const;

	// :1999-2049
	class PathIterator {
	public:
		PathIterator(const VariantGraph* g, Graph::vertex v, const std::string& sample);
		const Vertex* operator*() const { return cur; }
		void operator++();
		bool done() const { return is_done; }
	private:
		const VariantGraph* vg; const Vertex* cur; uint32_t s_id; bool is_done;
	};
	// :2092-2114
	class BfsIterator {
	public:
		BfsIterator(const VariantGraph* g, Graph::vertex v, uint64_t r) : vg(g), itr(&g->topology, v, r) {}
		const Vertex* operator*() const { return &vg->get_vertex(*itr); }
		void operator++() { ++itr; }
		bool done() const { return itr.done(); }
	private:
		const VariantGraph* vg; Graph::GraphIterator itr;
	};
	PathIterator find(uint64_t vertex_id, const std::string& sample) const { return PathIterator(this, (Graph::vertex)vertex_id, sample); }
	PathIterator find(const std::string& sample) const { return PathIterator(this, 0, sample); }
	BfsIterator find(Graph::vertex v = 0, uint64_t radius = UINT64_MAX) const { return BfsIterator(this, v, radius); }

	// state (public: the oracle is test infrastructure, tests introspect it)
	uint64_t seq_length = 0, num_vertices = 0, ref_length = 0, num_samples = 0;
	std::string chr, prefix;
	bool read_only = false, use_bit_vector = false;
	ConstructOpts opts;
	std::map<uint64_t, uint64_t> idx_vertex_id;
	std::unordered_map<uint32_t, std::string> idsample_map;
	std::unordered_map<uint64_t, uint32_t> sampleclass_map;
	std::unordered_map<std::string, uint32_t> sampleid_map;
	std::vector<Vertex> vertices;
	std::vector<uint8_t> seq_buffer;      // 3-bit codes, one per byte in memory
	BitVec sample_vector;
	Graph topology;
	// construct statistics (add_vcfs :634-637, :730-732)
	uint64_t num_vars = 0, num_mutations = 0, num_mutations_samples = 0;

private:
	void init_ref(const std::string& ref);
	enum MUT { INSERTION, DELETION, SUBSTITUTION };
	Vertex* create_vertex(uint64_t id, uint64_t offset, uint64_t length, uint32_t class_id,
	                      const std::vector<SampleInfo>& samples);                          // :1068-1113
	Vertex* add_vertex(const std::string& seq, uint64_t index, uint32_t class_id, const SampleStruct& s);  // :757-785
	void split_vertex(uint64_t vertex_id, uint64_t pos, Graph::vertex* new_vertex);           // :1115-1160
	void split_vertex(uint64_t vertex_id, uint64_t pos1, uint64_t pos2, Graph::vertex* n1, Graph::vertex* n2); // :1162-1167
	void add_sample_vector(const BitVec& v, uint64_t class_id);                               // :787-801
	uint32_t find_sample_vector_or_add(const std::vector<SampleStruct>& l);                   // :803-832
	bool update_vertex_sample_class(Graph::vertex id, const std::vector<SampleStruct>& l);    // :834-873
	uint32_t get_popcnt(uint32_t class_id) const;                                             // :1023-1041
	uint32_t get_sample_id(uint32_t class_id, uint32_t index) const;                          // :902-942
	void add_sample_to_vertex(Graph::vertex id, uint64_t sample_idx, const SampleStruct& s);  // :1453-1464
	void update_idx_vertex_id_map(const Vertex& v);                                           // :1280-1287
	void validate_ref_path_edge(Graph::vertex src, Graph::vertex dest) const;                 // :1483-1494
	void add_mutation(std::string ref, std::string alt, uint64_t pos, std::vector<SampleStruct>& l);  // :1509-1881
	void fix_sample_indexes();                                                                // :1883-1997
	bool is_bit_vector(const Vertex& v) const { return v.s_info[0].has_sid ? false : true; } // :1291-1294
	bool detect_encoding(const std::string& vcf_file);                                        // :568-617
};

// ---------------------------------------------------------------- Index  (index.h:27-50)
class Index {
public:
	explicit Index(const VariantGraph* vg);          // index.h:53-106
	explicit Index(const std::string& prefix);       // index.h:108-117
	Graph::vertex find(uint64_t pos) const;                          // :119-133
	Graph::vertex find(uint64_t pos, uint64_t& ref_node_rank) const; // :135-148
	bool is_empty(uint64_t x, uint64_t y) const;                     // :150-166
	Graph::vertex previous(uint64_t ref_node_rank) const;            // :168-172
	void serialize(const std::string& prefix) const;                 // :174-179

	uint64_t size_bits = 0;                 // rank_rrrb.size()
	std::vector<uint64_t> ones;             // sorted positions of set bits (rank/select support)
	std::vector<uint32_t> node_list;
	uint64_t rank(uint64_t pos) const;      // #ones in [0,pos)
};

// ---------------------------------------------------------------- operators (query.h)
struct Variant {                                   // query.h:30-36
	uint64_t var_pos = 0; bool var_pos_set = false;  // reference leaves var_pos uninitialised (:322)
	std::string ref, alt;
	std::vector<std::pair<std::string, std::string>> samples;
};
struct QueryLog { std::string out; std::string err; };   // stdout count lines / console->error text

Graph::vertex get_prev_vertex_with_sample(const VariantGraph* vg, const Index* idx, uint64_t pos,
                                          const std::string& sample_id, uint64_t& ref_pos,
                                          uint64_t& sample_pos, bool* ub = nullptr);     // :57-113
std::string query_sample_from_ref(const VariantGraph* vg, const Index* idx, uint64_t x, uint64_t y,
                                  const std::string& sample_id, bool print = false,
                                  const std::string& outfile = "", bool* ub = nullptr);   // :120-189
std::string query_sample_from_sample(const VariantGraph* vg, const Index* idx, uint64_t x, uint64_t y,
                                     const std::string& sample_id, bool print = false,
                                     const std::string& outfile = "", bool* ub = nullptr, bool* hang = nullptr);   // :195-261
bool get_samples(const Vertex* v, const VariantGraph* vg,
                 std::vector<std::pair<std::string, std::string>>& sample_ids);       // :268-285
bool next_variant_in_ref(const VariantGraph* vg, const Index* idx, uint64_t pos, std::vector<Variant>& vars,
                         uint64_t& next_pos, uint64_t end = UINT64_MAX);             // :297-436
bool closest_var(const VariantGraph* vg, const Index* idx, uint64_t pos, std::vector<Variant>& vars,
                 bool print = false, const std::string& outfile = "");               // :441-483
std::vector<Variant> get_sample_var_in_ref(const VariantGraph* vg, const Index* idx, uint64_t x, uint64_t y,
                                           const std::string& sample_id, bool print = false,
                                           const std::string& outfile = "", QueryLog* log = nullptr,
                                           bool* ub = nullptr);                      // :618-729
std::vector<Variant> get_sample_var_in_sample(const VariantGraph* vg, const Index* idx, uint64_t x, uint64_t y,
                                              const std::string& sample_id, bool print = false,
                                              const std::string& outfile = "", QueryLog* log = nullptr,
                                              bool* ub = nullptr, bool* hang = nullptr);   // :490-612
std::vector<Variant> get_var_in_ref(const VariantGraph* vg, const Index* idx, uint64_t x, uint64_t y,
                                    bool print = false, const std::string& outfile = "",
                                    QueryLog* log = nullptr);                        // :736-784
std::vector<std::pair<std::string, std::string>>
samples_has_var(const VariantGraph* vg, const Index* idx, uint64_t pos, const std::string& ref,
                const std::string& alt, bool print = false, const std::string& outfile = "",
                QueryLog* log = nullptr);                                            // :792-823

// commands.cc:64-93 / :96-111
std::vector<std::pair<uint64_t, uint64_t>> read_regions(std::string region);
std::vector<std::string> read_sequences(std::string s);

// ---------------------------------------------------------------- codecs (ser/ layout)
namespace codec {
bool write_int_vector0(const std::string& path, const std::vector<uint64_t>& vals, uint8_t width);
bool read_int_vector0(const std::string& path, std::vector<uint64_t>& vals, uint8_t* width = nullptr);
bool write_int_vector32(const std::string& path, const std::vector<uint32_t>& vals);
bool read_int_vector32(const std::string& path, std::vector<uint32_t>& vals);
bool write_rrr127(const std::string& path, const BitVec& bv);
bool read_rrr127(const std::string& path, BitVec& bv);
bool write_vertex_block(const std::string& path, const Vertex* v, size_t n, int gzip_level);
bool read_vertex_block(const std::string& path, std::vector<Vertex>& out);
void encode_vertex_list(const Vertex* v, size_t n, std::string& out);     // raw VariantGraphVertexList bytes
bool decode_vertex_list(const uint8_t* p, size_t len, std::vector<Vertex>& out);
uint8_t bits_needed(uint64_t maxval);    // sdsl::util::bit_compress width
}

uint64_t murmur_hash_64a(const void* key, int len, unsigned int seed);
char map_int(uint8_t base);      // util.cc:32-41
uint8_t map_base(char base);     // util.cc:44-53
void read_fasta(const std::string& fasta_file, std::string& chr, std::string& ref);  // util.cc:82-105

}  // namespace vso
