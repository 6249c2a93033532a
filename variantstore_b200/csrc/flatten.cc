// libvsgpu host side — flattener.  See flatten.h.
#include "flatten.h"

#include <algorithm>
#include <cstring>
#include <stdexcept>

namespace vsgpu {
namespace {
[[noreturn]] void fail(const std::string& m) { throw std::runtime_error("vsgpu: " + m); }
}

uint64_t hash_ref_alt(const char* ref, size_t nref, const char* alt, size_t nalt) {
	uint64_t h = 14695981039346656037ULL;
	for (size_t i = 0; i < nref; i++) { h ^= (unsigned char)ref[i]; h *= 1099511628211ULL; }
	h ^= 0xFF; h *= 1099511628211ULL;
	for (size_t i = 0; i < nalt; i++) { h ^= (unsigned char)alt[i]; h *= 1099511628211ULL; }
	return h;
}

void flatten(const SerData& d, FlatIndex& f) {
	PhaseClock pc;
	const uint32_t nv = d.num_vertices;
	const uint32_t ns = d.num_samples;
	f.ref_length = d.ref_length; f.index_bits = d.index_bits; f.num_samples = ns; f.class_mode = d.class_mode;

	// ------------------------------------------------------------ carrier sets
	// class mode: row c of `bitmap` = sample class c (variant_graph.h:787-801 lays class c>=1 at bits
	// [(c-1)*num_samples, c*num_samples) of one long vector); row 0 = {ref}.
	// explicit-id mode: one id list per vertex that lists a non-ref sample; set 0 = {}.
	std::vector<uint32_t> vset(nv, 0);        // carrier-set id per vertex
	std::vector<uint8_t> vhas_ref(nv, 0), vhas_nonref(nv, 0);
	std::vector<uint32_t> vref_index(nv, 0), vfirst_nonref(nv, 0);
	if (d.class_mode) {
		const uint64_t C = d.sample_vector_bits / ns;
		f.words_per_set = (ns + 63) / 64; f.num_sets = (uint32_t)C + 1;
		f.bitmap.assign((uint64_t)f.num_sets * f.words_per_set, 0);
		f.bitmap[0] = 1;
		auto get = [&](uint64_t pos, unsigned len) {
			uint64_t wi = pos >> 6; unsigned off = (unsigned)(pos & 63);
			uint64_t v = d.sample_vector[wi] >> off;
			if (off + len > 64) v |= d.sample_vector[wi + 1] << (64 - off);
			return len < 64 ? v & ((1ULL << len) - 1) : v;
		};
		for (uint64_t c = 1; c <= C; c++)
			for (uint32_t w = 0; w < f.words_per_set; w++) {
				unsigned len = std::min<uint32_t>(64, ns - w * 64);
				f.bitmap[c * f.words_per_set + w] = get((c - 1) * ns + (uint64_t)w * 64, len);
			}
		std::vector<uint32_t> set_pop(f.num_sets, 0), set_first_nonref(f.num_sets, 0);
		for (uint32_t c = 0; c < f.num_sets; c++) {
			uint32_t pc = 0, first = 0;
			for (uint32_t w = 0; w < f.words_per_set; w++) {
				uint64_t x = f.bitmap[(uint64_t)c * f.words_per_set + w];
				pc += (uint32_t)__builtin_popcountll(x);
				if (w == 0) x &= ~1ULL;
				if (!first && x) first = w * 64 + (uint32_t)__builtin_ctzll(x);
			}
			set_pop[c] = pc; set_first_nonref[c] = first;
		}
		for (uint32_t v = 0; v < nv; v++) {
			uint32_t c = d.v_class[v];
			if (c >= f.num_sets) fail("vertex " + std::to_string(v) + " names sample class " + std::to_string(c) + " beyond sample_vector.sdsl");
			vset[v] = c;
			vhas_ref[v] = (uint8_t)(f.bitmap[(uint64_t)c * f.words_per_set] & 1);
			vhas_nonref[v] = set_pop[c] > vhas_ref[v];
			vfirst_nonref[v] = set_first_nonref[c];
			// s_info[i] belongs to the i-th set bit (variant_graph.h:1302-1315); ref is bit 0 -> entry 0
			if (vhas_ref[v] && d.v_sinfo_begin[v] < d.v_sinfo_begin[v + 1]) vref_index[v] = d.v_first_index[v];
		}
	} else {
		f.words_per_set = 0;
		f.list_begin.push_back(0); f.list_begin.push_back(0);   // set 0 = {}
		uint32_t next = 1;
		for (uint32_t v = 0; v < nv; v++) {
			bool nonref = false;
			for (uint64_t i = d.v_sinfo_begin[v]; i < d.v_sinfo_begin[v + 1]; i++) {
				uint32_t id = d.s_sample_id[i];
				if (id == 0) { if (!vhas_ref[v]) { vhas_ref[v] = 1; vref_index[v] = d.v_ref0_index[v]; } }
				else { if (!nonref) { nonref = true; vfirst_nonref[v] = id; } f.list_ids.push_back(id); }
			}
			if (nonref) { vhas_nonref[v] = 1; vset[v] = next++; f.list_begin.push_back(f.list_ids.size()); }
		}
		f.num_sets = next;
	}
	auto carries = [&](uint32_t v, uint32_t s) { return s == 0 ? (bool)vhas_ref[v] : f.member(s, vset[v]); };

	auto out_begin = [&](uint32_t v) { return d.adj_begin[v]; };
	auto out_end = [&](uint32_t v) { return d.adj_begin[v + 1]; };

	// VariantGraph::get_neighbor_vertex (variant_graph.h:1402-1451): first out-neighbour carrying
	// `sid` (sid != 0), else the ref-carrying neighbour with the smallest index (first wins on ties).
	auto neighbor_vertex = [&](uint32_t v, uint32_t sid) -> uint32_t {
		uint32_t next = 0, min_idx = UINT32_MAX;
		for (uint64_t i = out_begin(v); i < out_end(v); i++) {
			uint32_t n = d.adj[i];
			if (n == UINT32_MAX) continue;
			if (vhas_ref[n] && min_idx > vref_index[n]) { next = n; min_idx = vref_index[n]; }
			if (sid != 0 && carries(n, sid)) return n;
		}
		return next;
	};

	pc.lap("flatten: carrier sets");
	// ------------------------------------------------------------ backbone
	f.vertex_bb.assign(nv, kNone);
	for (uint32_t v = 0, steps = 0;; steps++) {
		if (steps > nv) fail("ref path does not terminate (cycle)");
		if (!vhas_ref[v]) fail("ref path reaches vertex " + std::to_string(v) + " which does not carry ref");
		f.vertex_bb[v] = (uint32_t)f.bb_vertex.size();
		f.bb_vertex.push_back(v); f.vstart.push_back(vref_index[v]); f.vlen.push_back(d.v_length[v]);
		uint32_t n = neighbor_vertex(v, 0);
		if (n == 0) break;
		v = n;
	}
	const uint32_t M = f.M = (uint32_t)f.bb_vertex.size();
	if (M >= kEntTgtMask) fail("backbone too long for one shard");
	for (uint32_t k = 0; k + 1 < M; k++)
		if ((uint64_t)f.vstart[k] + f.vlen[k] != f.vstart[k + 1]) fail("backbone is not a gap-free tiling at vertex " + std::to_string(f.bb_vertex[k]));
	f.bb_set.assign(M, 0);
	for (uint32_t k = 0; k < M; k++) if (vhas_nonref[f.bb_vertex[k]]) f.bb_set[k] = vset[f.bb_vertex[k]];

	pc.lap("flatten: backbone");
	// ------------------------------------------------------------ distinct starts <-> loaded index
	const uint32_t D = f.D = (uint32_t)d.index_ones.size();
	if (D == 0 || d.node_list.size() != D) fail("index.sdsl and ref_node_id.sdsl disagree");
	f.dstart.resize(D); f.dlev.assign(D + 1, DLevel{M, 0, 0, 0});
	for (uint32_t i = 0; i < D; i++) {
		f.dstart[i] = d.index_ones[i] + 1;
		uint32_t v = d.node_list[i];
		uint32_t k = v < nv ? f.vertex_bb[v] : kNone;
		if (k == kNone || f.vstart[k] != f.dstart[i] || (k > 0 && f.vstart[k - 1] == f.vstart[k]))
			fail("position index does not match the graph's ref path at entry " + std::to_string(i));
		f.dlev[i].k = k;
	}
	{ uint32_t distinct = 0; for (uint32_t k = 0; k < M; k++) if (k == 0 || f.vstart[k] != f.vstart[k - 1]) distinct++; if (distinct != D) fail("position index misses backbone starts"); }

	pc.lap("flatten: distinct starts");
	// ------------------------------------------------------------ per backbone vertex: records + entries
	auto seq_eq = [&](uint32_t va, uint32_t vb) {   // sequences of two vertices (kNone = "")
		uint32_t la = va == kNone ? 0 : d.v_length[va], lb = vb == kNone ? 0 : d.v_length[vb];
		if (la != lb) return false;
		if (la == 0) return true;
		return memcmp(&d.seq[d.v_offset[va]], &d.seq[d.v_offset[vb]], la) == 0;
	};
	static const char kBase[8] = {'A', 'C', 'T', 'G', 'N', 5, 5, 5};   // util.cc:32-41
	auto seq_str = [&](uint32_t v, std::string& s) { s.clear(); if (v == kNone) return; for (uint32_t i = 0; i < d.v_length[v]; i++) s += kBase[d.seq[d.v_offset[v] + i] & 7]; };

	f.rec_begin.assign(M + 1, 0); f.cent_begin.assign(M + 1, 0);
	std::vector<uint32_t> outdeg(M, 0), ncar(M, 0);
	std::string sref, salt;
	for (uint32_t k = 0; k < M; k++) {
		const uint32_t v = f.bb_vertex[k];
		const uint32_t succ = k + 1 < M ? f.bb_vertex[k + 1] : 0;     // *next_it; vertex 0 once the path iterator is done
		f.rec_begin[k] = (uint32_t)f.rec_k.size(); f.cent_begin[k] = (uint32_t)f.cent.size();
		// next_ref_pos computed at P[k] (query.h:660-674): last ref-carrying neighbour wins
		uint32_t nrp = f.vstart[k] + f.vlen[k], nref = kNone;
		for (uint64_t i = out_begin(v); i < out_end(v); i++) { uint32_t n = d.adj[i]; if (n != UINT32_MAX && vhas_ref[n]) { nrp = vref_index[n]; nref = n; } }
		f.bb_nrp.push_back(nrp); f.bb_nref.push_back(nref);
		{   // t2: the first ref-carrying neighbour sets next_ref_pos and the scan stops (query.h:143-151)
			uint32_t n1 = f.vstart[k] + f.vlen[k];
			for (uint64_t i = out_begin(v); i < out_end(v); i++) { uint32_t n = d.adj[i]; if (n != UINT32_MAX && vhas_ref[n]) { n1 = vref_index[n]; break; } }
			f.nrp1.push_back(n1);
			if (f.vlen[k] && d.v_offset[v] != (uint64_t)f.vstart[k] - 1 && f.t2_ok) { f.t2_ok = false; f.t2_why = "sequence of backbone vertex " + std::to_string(v) + " is not at offset start - 1 of seq_buffer"; }
			if (k + 1 < M && n1 < f.vstart[k + 1] && f.t2_ok) { f.t2_ok = false; f.t2_why = "backbone vertex " + std::to_string(v) + " has a ref neighbour behind its successor"; }
		}
		const size_t rec0 = f.rec_k.size();
		for (uint64_t i = out_begin(v); i < out_end(v); i++) {
			const uint32_t n = d.adj[i];
			if (n == UINT32_MAX) continue;
			outdeg[k]++;
			const bool n_bb = f.vertex_bb[n] != kNone;
			if (!n_bb) {
				uint32_t deg = 0, nn = 0;
				for (uint64_t j = out_begin(n); j < out_end(n); j++) if (d.adj[j] != UINT32_MAX) { deg++; nn = d.adj[j]; }
				if (deg > 1) fail("Sample vertex has more than 1 neighbor: " + std::to_string(n));
				if (deg == 1 && f.vertex_bb[nn] == kNone) fail("consecutive mutation: alt vertex " + std::to_string(n) + " does not rejoin the ref path");
			}
			// ---- branch record (query.h:322-415)
			if (n != succ) {
				uint32_t pos = 0, refv = kNone, altv = kNone; uint8_t flags = 0;
				if (vhas_nonref[n]) {
					if (vhas_ref[n]) {                          // deletion (:336-350)
						refv = succ; pos = vref_index[succ];
					} else {
						uint32_t prev_ref_idx = f.vstart[k];
						uint32_t step = neighbor_vertex(n, vfirst_nonref[n]);          // ++dfs_it (:356-358)
						uint32_t next_ref_idx = vhas_ref[step] ? vref_index[step] : prev_ref_idx;   // `sample` keeps P[k]'s entry on failure
						if (next_ref_idx == prev_ref_idx + f.vlen[k]) { altv = n; pos = next_ref_idx - 1; }   // insertion (:369-376)
						else { altv = n; refv = succ; pos = vref_index[succ]; }        // substitution (:378-392)
					}
				} else flags |= 4;                            // no non-ref sample: the reference pushes an uninitialised Variant
				f.rec_k.push_back(k); f.rec_vertex.push_back(n); f.rec_pos.push_back(pos); f.rec_refv.push_back(refv); f.rec_altv.push_back(altv);
				f.rec_flags.push_back(flags);
			}
			// ---- compact entry for the sample walk
			if (vhas_nonref[n]) {
				CEntry e; e.src = k; e.set_id = vset[n]; e.arrival = nrp;
				if (n_bb) e.tgt = f.vertex_bb[n];
				else {
					uint32_t nn = kNone;
					for (uint64_t j = out_begin(n); j < out_end(n); j++) if (d.adj[j] != UINT32_MAX) nn = d.adj[j];
					if (nn == kNone) e.tgt = kEntAlt | kEntTgtNone;
					else { uint32_t tk = f.vertex_bb[nn]; e.tgt = kEntAlt | tk | (f.bb_set[tk] ? kEntTgtCarriers : 0); }
				}
				f.cent.push_back(e); f.cent_vertex.push_back(n); ncar[k]++;
				if (n_bb) { f.cent_seq.push_back(0); f.cent_seq.push_back(0); }
				else { f.cent_seq.push_back(d.v_offset[n]); f.cent_seq.push_back(d.v_length[n]); }
			}
		}
		if (k + 1 < M && nrp != f.vstart[k + 1]) {        // arrival at P[k+1] along the backbone is out of step
			f.cent.push_back(CEntry{k, kEntMarker | (k + 1), 0, nrp}); f.cent_vertex.push_back(kNone);
			f.cent_seq.push_back(0); f.cent_seq.push_back(0);
		}
		// ---- t7: which records survive the dedup of a fresh next_variant_in_ref call (:397-414)
		std::vector<size_t> kept;
		for (size_t r = rec0; r < f.rec_k.size(); r++) {
			bool push = true;
			if (!kept.empty()) {
				size_t b = kept.back();
				if (f.rec_pos[b] == f.rec_pos[r] && seq_eq(f.rec_altv[b], f.rec_altv[r])) push = false;
				else if (kept.size() > 1 && f.rec_pos[b] == f.rec_pos[r]) {
					for (size_t q = kept.size(); q-- > 0;) {
						size_t o = kept[q];
						if (f.rec_pos[o] < f.rec_pos[r]) break;
						if (f.rec_pos[o] == f.rec_pos[r] && seq_eq(f.rec_altv[o], f.rec_altv[r])) { push = false; break; }
					}
				}
			}
			if (push) { kept.push_back(r); f.rec_flags[r] |= 1; }
		}
	}
	f.rec_begin[M] = f.R = (uint32_t)f.rec_k.size(); f.cent_begin[M] = (uint32_t)f.cent.size();
	if (f.cent.size() >= (1u << 30)) fail("too many walk entries for one shard");
	f.row_words = (uint32_t)(((f.cent.size() + 32) / 32 + 31) / 32 * 32);   // one spare word past the end
	f.marker_bits.assign(f.row_words, 0);
	for (size_t c = 0; c < f.cent.size(); c++) if (f.cent[c].tgt & kEntMarker) f.marker_bits[c >> 5] |= 1u << (c & 31);

	pc.lap("flatten: records + entries");
	// suspect duplicates for t6: an earlier record with the same (pos, alt) inside the preceding run
	// of records whose pos is >= this one's.  Ranges containing a suspect are re-counted on the host
	// with the literal rule.
	f.rec_dup_prefix.assign(f.R + 1, 0);
	for (uint32_t r = 0; r < f.R; r++) {
		bool suspect = false;
		for (uint32_t q = r; q-- > 0;) {
			if (f.rec_pos[q] < f.rec_pos[r]) break;
			if (f.rec_pos[q] == f.rec_pos[r] && seq_eq(f.rec_altv[q], f.rec_altv[r])) { suspect = true; break; }
		}
		if (suspect) { f.rec_flags[r] |= 2; f.has_suspect_dups = true; }
		f.rec_dup_prefix[r + 1] = f.rec_dup_prefix[r] + (suspect ? 1 : 0);
	}
	f.rec_hash.resize(f.R);
	for (uint32_t r = 0; r < f.R; r++) { seq_str(f.rec_refv[r], sref); seq_str(f.rec_altv[r], salt); f.rec_hash[r] = hash_ref_alt(sref.data(), sref.size(), salt.data(), salt.size()); }

	pc.lap("flatten: suspect dups");
	// ------------------------------------------------------------ distinct-start level tables
	f.dinfo.assign(D, 0); f.t7_lo.assign(D, 0); f.t7_hi.assign(D, 0);
	std::vector<uint32_t> next_branchy(M + 1, M);
	for (uint32_t k = M; k-- > 0;) next_branchy[k] = f.rec_begin[k + 1] > f.rec_begin[k] ? k : next_branchy[k + 1];
	for (uint32_t i = 0; i < D; i++) {
		const uint32_t k = f.dlev[i].k;
		f.dlev[i].rec_lo = f.rec_begin[k];
		f.dlev[i].rec_hi_prev = k ? f.rec_begin[k - 1] : 0;
		f.dlev[i].cent_begin = f.cent_begin[k];
		if (ncar[k] > 0xFFFF || outdeg[k] > 0xFFFF) fail("out-degree too large at backbone vertex " + std::to_string(f.bb_vertex[k]));
		f.dinfo[i] = (uint64_t)f.cent_begin[k] | ((uint64_t)ncar[k] << 32) | ((uint64_t)outdeg[k] << 48);
		uint32_t kb = next_branchy[k];
		if (kb < M) { f.t7_lo[i] = f.rec_begin[kb]; f.t7_hi[i] = f.rec_begin[kb + 1]; }
	}
	f.dlev[D] = DLevel{M, f.R, M ? f.rec_begin[M - 1] : 0, (uint32_t)f.cent.size()};

	// t2: first backbone index whose nrp1 reaches each distinct start (prefix maxima are monotone)
	f.first_reach.assign(D + 1, M);
	{
		uint32_t j = 0, pm = 0;     // pm = max nrp1[0..j)
		for (uint32_t i = 0; i <= D; i++) {
			while (j < M && (i < D ? std::max(pm, f.nrp1[j]) < f.dstart[i] : std::max(pm, f.nrp1[j]) <= f.dstart[D - 1])) { pm = std::max(pm, f.nrp1[j]); j++; }
			f.first_reach[i] = j;     // first j with nrp1[j] >= dstart[i]  (M: none)
		}
	}
	pc.lap("flatten: level tables");
	// ------------------------------------------------------------ back-walk forest
	// State c (= cur_ref_node_idx) examines node_list[c-1] and moves to c - outdeg; states 0 and 1 end
	// the walk at vertex 0.  parent < child, so subtree sizes and pre-order times need no recursion.
	{
		std::vector<uint32_t> parent(D + 1, 0), sz(D + 1, 1), slot(D + 1, 0);
		for (uint32_t c = 2; c <= D; c++) { uint32_t deg = outdeg[f.dlev[c - 1].k]; parent[c] = c >= deg ? c - deg : 0; if (parent[c] >= c) parent[c] = c - 1; }
		for (uint32_t c = D; c >= 2; c--) sz[parent[c]] += sz[c];
		f.dtin.assign(D + 1, 0);
		f.dtin[0] = 0; slot[0] = 1;
		if (D >= 1) { f.dtin[1] = sz[0]; slot[1] = f.dtin[1] + 1; }
		for (uint32_t c = 2; c <= D; c++) { f.dtin[c] = slot[parent[c]]; slot[parent[c]] += sz[c]; slot[c] = f.dtin[c] + 1; }
		f.cent_anc.assign(2 * f.cent.size(), 0);
		std::vector<uint32_t> k2c(M, 0);
		for (uint32_t i = 1; i < D; i++) k2c[f.dlev[i].k] = i + 1;        // node_list[i] is examined by state i + 1 (>= 2)
		for (size_t e = 0; e < f.cent.size(); e++) {
			uint32_t c = (f.cent[e].tgt & kEntMarker) ? 0 : k2c[f.cent[e].src];
			if (c >= 2) { f.cent_anc[2 * e] = f.dtin[c]; f.cent_anc[2 * e + 1] = f.dtin[c] + sz[c] - 1; }
			else { f.cent_anc[2 * e] = 1; f.cent_anc[2 * e + 1] = 0; }
		}
	}
	pc.lap("flatten: back-walk forest");
}

}  // namespace vsgpu
