// libvsgpu host side — reader for VariantStore's serialised index directory (`ser/`).
//
// Replaces the load half of the reference:
//   Index::Index(prefix)                 include/index.h:108-117
//   Graph::Graph(prefix)                 include/graph.h:149-172   (+ qf_deserialize gqf_file.c:274-332)
//   VariantGraph::VariantGraph(prefix,…) include/variant_graph.h:366-446 (+ stream::for_each stream.hpp:71-112)
// Everything is decoded once, straight into structure-of-arrays; nothing is paged back in later.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace vsgpu {

struct SerData {
	// sampleid_map.lst (variant_graph.h:541-556)
	std::string chr;
	uint64_t ref_length = 0;
	uint32_t num_samples = 0;                   // including "ref" (id 0)
	std::vector<std::string> sample_names;      // by id

	// index.sdsl / ref_node_id.sdsl (index.h:174-179)
	uint64_t index_bits = 0;                    // rank_rrrb.size()
	std::vector<uint32_t> index_ones;           // positions of set bits, ascending (= vertex start - 1)
	std::vector<uint32_t> node_list;

	// vertex_list_<k>.proto (variantgraphvertex.proto)
	uint32_t num_vertices = 0;
	bool class_mode = true;                     // bit-vector encoding (sampleclass_id present, no sample_id)
	std::vector<uint32_t> v_offset, v_length, v_class;
	std::vector<uint64_t> v_sinfo_begin;        // CSR into the s_info arrays, num_vertices + 1
	std::vector<uint32_t> v_first_index;        // sample_info.index of the vertex's first s_info (the ref entry in class mode)
	std::vector<uint32_t> v_ref0_index;         // ... of its first s_info that names sample id 0 (explicit-id mode)
	std::vector<uint32_t> s_sample_id;          // explicit-id mode only
	std::vector<uint8_t> s_flags;               // bit0 phase, bit1 gt_1, bit2 gt_2

	// seq_buffer.sdsl: one 3-bit base code per byte (util.h:44: A0 C1 T2 G3 N4)
	std::vector<uint8_t> seq;

	// sample_vector.sdsl: concatenated class bitmaps, class c>=1 at bits [(c-1)*num_samples, c*num_samples)
	uint64_t sample_vector_bits = 0;
	std::vector<uint64_t> sample_vector;

	// adj_list.cqf + aux_vertex_list*.sdsl: out-neighbour lists in the iteration order the
	// reference observes after loading (std::unordered_set rebuilt by in-order insertion).
	std::vector<uint64_t> adj_begin;            // num_vertices + 1
	std::vector<uint32_t> adj;
	uint64_t cqf_distinct = 0;                  // ndistinct_elts = "#Vertices" the CLI prints
};

// Throws std::runtime_error with a message naming the file that failed.
void load_ser(const std::string& prefix, SerData& out);
// Throws when a vertex names a slice of seq_buffer.sdsl that runs past its end (load_ser and the index cache call it).
void check_seq_ranges(const SerData& d);

// sample_info.index of every s_info in SerData order (second pass over vertex_list_<k>.proto); `expect`
// = v_sinfo_begin.back()
void load_sample_indexes(const std::string& prefix, uint64_t expect, std::vector<uint32_t>& out);

// VSGPU_TRACE=1: wall-clock seconds of the phases of vsgpu_open on stderr
struct PhaseClock {
	bool on; double t0;
	PhaseClock();
	void lap(const char* what);
};

}  // namespace vsgpu
