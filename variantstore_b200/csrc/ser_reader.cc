// libvsgpu host side — `ser/` reader.  See ser_reader.h for the reference entry points replaced.
//
// On-disk formats handled here:
//   * sdsl-lite containers — int_vector<0> (u64 bit size, u8 width, words), int_vector<32> /
//     bit_vector (u64 bit size, words), rrr_vector<127> (size, bt, btnr, btnrp, rank, invert).
//     sdsl-lite is not vendored by the reference; layouts follow the published sdsl-lite v2.1
//     serialisers.  The rrr layout has no golden file anywhere in the reference ("parity unpinned").
//   * gzip-framed protobuf vertex blocks (stream.hpp:25-52, variantgraphvertex.proto:6-26).
//   * the Counting Quotient Filter image (gqf_int.h:37-101, gqf_file.c:259-272).
#include "ser_reader.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <fstream>
#include <stdexcept>
#include <sys/mman.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <unordered_set>
#include <zlib.h>

namespace vsgpu {
namespace {

[[noreturn]] void fail(const std::string& m) { throw std::runtime_error("vsgpu: " + m); }

// ------------------------------------------------------------------ read-only file mapping
struct Mapped {
	const uint8_t* p = nullptr; size_t n = 0; int fd = -1;
	explicit Mapped(const std::string& path) {
		fd = open(path.c_str(), O_RDONLY);
		if (fd < 0) fail("cannot open " + path);
		struct stat st;
		if (fstat(fd, &st) != 0) { close(fd); fail("cannot stat " + path); }
		n = (size_t)st.st_size;
		if (n) { void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0); if (m == MAP_FAILED) { close(fd); fail("cannot map " + path); } p = (const uint8_t*)m; }
	}
	~Mapped() { if (p) munmap((void*)p, n); if (fd >= 0) close(fd); }
};

struct Cursor {
	const uint8_t* p; const uint8_t* end; const std::string& name;
	uint64_t u64() { if (end - p < 8) fail("truncated " + name); uint64_t v; memcpy(&v, p, 8); p += 8; return v; }
	uint8_t u8() { if (end - p < 1) fail("truncated " + name); return *p++; }
	const uint64_t* words(uint64_t n) { if ((uint64_t)(end - p) < n * 8) fail("truncated " + name); const uint64_t* w = (const uint64_t*)p; p += n * 8; return w; }
};

// bit field reader over little-endian u64 words (unaligned word pointer tolerated via memcpy)
inline uint64_t bits_at(const uint64_t* w, uint64_t nwords, uint64_t pos, unsigned len) {
	if (len == 0) return 0;
	uint64_t wi = pos >> 6; unsigned off = (unsigned)(pos & 63);
	uint64_t a; memcpy(&a, w + wi, 8);
	uint64_t v = a >> off;
	if (off + len > 64 && wi + 1 < nwords) { uint64_t b; memcpy(&b, w + wi + 1, 8); v |= b << (64 - off); }
	if (len < 64) v &= (1ULL << len) - 1;
	return v;
}

struct PackedInts { const uint64_t* w = nullptr; uint64_t nwords = 0, count = 0; unsigned width = 0;
	uint64_t operator[](uint64_t i) const { return bits_at(w, nwords, i * width, width); } };

PackedInts read_iv0(Cursor& c) {
	PackedInts r; uint64_t bits = c.u64(); r.width = c.u8();
	if (r.width == 0 || r.width > 64) fail("bad int_vector width in " + c.name);
	r.nwords = (bits + 63) / 64; r.w = c.words(r.nwords); r.count = bits / r.width;
	return r;
}
struct PackedBits { const uint64_t* w = nullptr; uint64_t nwords = 0, nbits = 0;
	bool operator[](uint64_t i) const { uint64_t a; memcpy(&a, w + (i >> 6), 8); return (a >> (i & 63)) & 1; } };
PackedBits read_bv(Cursor& c) { PackedBits r; r.nbits = c.u64(); r.nwords = (r.nbits + 63) / 64; r.w = c.words(r.nwords); return r; }

// ------------------------------------------------------------------ rrr_vector<127> decoder
// Block types are popcounts (7 bits); each block's bit pattern is stored as its index in the
// combinatorial number system among the C(127, popcount) patterns, enumerated from bit 0;
// a superblock (32 blocks) may be stored complemented.
typedef unsigned __int128 u128;
struct RrrTables {
	u128 binom[128][128];
	uint8_t width[128];
	RrrTables() {
		memset(binom, 0, sizeof binom);
		for (int n = 0; n < 128; n++) { binom[n][0] = 1; for (int k = 1; k <= n; k++) binom[n][k] = binom[n - 1][k - 1] + binom[n - 1][k]; }
		for (int k = 0; k < 128; k++) { u128 c = binom[127][k]; int b = 0; if (c != 1) { while (c) { b++; c >>= 1; } } width[k] = (uint8_t)b; }
	}
};
const RrrTables& rrr_tables() { static RrrTables t; return t; }

// Decodes the whole vector into 64-bit words (bit i = word i/64, bit i%64); returns its length in
// bits.  Blocks are independent once the start of their offset field is known, so a first pass sums
// the field widths up to every chunk boundary (the file's own pointer samples are not trusted) and
// the chunks — whole numbers of 127-word spans, so no two threads share an output word — are
// decoded in parallel.
uint64_t decode_rrr127(const std::string& path, std::vector<uint64_t>& out) {
	Mapped m(path);
	Cursor c{m.p, m.p + m.n, path};
	const RrrTables& T = rrr_tables();
	const uint64_t size = c.u64();
	const PackedInts bt = read_iv0(c);
	const PackedBits btnr = read_bv(c);
	PackedInts btnrp = read_iv0(c); (void)btnrp;
	PackedInts rank = read_iv0(c); (void)rank;
	const PackedBits inv = read_bv(c);
	const uint64_t nblocks = (size + 126) / 127;
	if (nblocks > bt.count) fail("rrr block table too short in " + path);
	out.assign((size + 63) / 64 + 1, 0);
	constexpr uint64_t kChunk = 64 * 64;                       // blocks per chunk: 64 superblock pairs = 8128 words
	const uint64_t nchunks = (nblocks + kChunk - 1) / kChunk;
	std::vector<uint64_t> chunk_off(nchunks + 1, 0);
	{
		uint64_t off = 0;
		for (uint64_t blk = 0; blk < nblocks; blk++) {
			if (blk % kChunk == 0) chunk_off[blk / kChunk] = off;
			const unsigned k = (unsigned)bt[blk];
			if (k > 127) fail("bad rrr block type in " + path);
			off += T.width[k];
		}
		chunk_off[nchunks] = off;
		if (off > btnr.nbits) fail("rrr offset stream too short in " + path);
	}
	auto work = [&](uint64_t c0, uint64_t c1) {
		for (uint64_t ch = c0; ch < c1; ch++) {
			uint64_t off = chunk_off[ch];
			const uint64_t b1 = std::min(nblocks, (ch + 1) * kChunk);
			for (uint64_t blk = ch * kChunk; blk < b1; blk++) {
				const uint64_t base = blk * 127;
				const unsigned k = (unsigned)bt[blk], w = T.width[k];
				u128 nr = 0;
				if (w) { nr = bits_at(btnr.w, btnr.nwords, off, w > 64 ? 64 : w); if (w > 64) nr |= (u128)bits_at(btnr.w, btnr.nwords, off + 64, w - 64) << 64; }
				off += w;
				const bool flip = (blk / 32) < inv.nbits && inv[blk / 32];
				const unsigned len = (unsigned)std::min<uint64_t>(127, size - base);
				// unrank: walk the positions, deciding each bit by comparing against C(remaining - 1, left)
				u128 pat = 0;
				if (k == 127) pat = ~(u128)0 >> 1;
				else for (unsigned p = 0, left = k; left && p < 127; p++) { const u128 cnt = T.binom[126 - p][left]; if (nr >= cnt) { nr -= cnt; left--; pat |= (u128)1 << p; } }
				if (flip) pat = ~pat;
				if (len < 127) pat &= ((u128)1 << len) - 1; else pat &= ~(u128)0 >> 1;
				if (!pat) continue;
				// a word is only touched when it receives bits: the neighbouring chunk owns the words after ours
				const uint64_t wi = base >> 6; const unsigned sh = (unsigned)(base & 63);
				const uint64_t lo = (uint64_t)pat, hi = (uint64_t)(pat >> 64);
				const uint64_t w0 = lo << sh, w1 = (sh ? lo >> (64 - sh) : 0) | (hi << sh), w2 = sh ? hi >> (64 - sh) : 0;
				if (w0) out[wi] |= w0;
				if (w1) out[wi + 1] |= w1;
				if (w2) out[wi + 2] |= w2;
			}
		}
	};
	unsigned nt = (unsigned)std::min<uint64_t>(std::max(1u, std::thread::hardware_concurrency()), nchunks);
	if (nt <= 1) work(0, nchunks);
	else {
		std::vector<std::thread> th;
		const uint64_t per = (nchunks + nt - 1) / nt;
		for (unsigned t = 0; t < nt; t++) { const uint64_t a0 = t * per, a1 = std::min(nchunks, a0 + per); if (a0 < a1) th.emplace_back(work, a0, a1); }
		for (auto& t : th) t.join();
	}
	return size;
}

// ------------------------------------------------------------------ protobuf vertex blocks
inline bool varint(const uint8_t*& p, const uint8_t* e, uint64_t& v) {
	v = 0;
	for (unsigned s = 0; s < 64 && p < e; s += 7) { uint8_t b = *p++; v |= (uint64_t)(b & 0x7f) << s; if (b < 0x80) return true; }
	return false;
}
inline bool skip(const uint8_t*& p, const uint8_t* e, unsigned wt) {
	uint64_t t;
	if (wt == 0) return varint(p, e, t);
	if (wt == 1) { if (e - p < 8) return false; p += 8; return true; }
	if (wt == 2) { if (!varint(p, e, t) || (uint64_t)(e - p) < t) return false; p += t; return true; }
	if (wt == 5) { if (e - p < 4) return false; p += 4; return true; }
	return false;
}

// One decoded block.  Per s_info only the 3 phasing flags are kept (plus the sample id in
// explicit-id mode); of the `index` fields only the two per vertex that flatten() reads.
struct BlockSoA {
	std::vector<uint32_t> id, offset, length, cls, first_index, ref0_index; std::vector<uint8_t> has_cls;
	std::vector<uint64_t> sbegin; std::vector<uint32_t> ssid; std::vector<uint8_t> sflags;
	bool any_sid = false, any_cls = false;
	bool keep_index = false; std::vector<uint32_t> sindex;    // sample_info.index of every s_info (only when asked for)
};

void parse_sinfo(const uint8_t* p, const uint8_t* e, BlockSoA& b, const std::string& name, uint32_t& first_index, uint32_t& ref0_index, bool& seen_first, bool& seen_ref0) {
	uint32_t index = 0, sid = 0; uint8_t flags = 0, has = 0;
	while (p < e) {
		uint64_t tag, v;
		if (!varint(p, e, tag)) fail("bad s_info in " + name);
		unsigned f = (unsigned)(tag >> 3), wt = (unsigned)(tag & 7);
		if (wt == 0 && f >= 1 && f <= 5) {
			if (!varint(p, e, v)) fail("bad s_info in " + name);
			if (f == 1) index = (uint32_t)v; else if (f == 2) { if (!has) { has = 1; sid = (uint32_t)v; } }
			else if (v) flags |= (uint8_t)(1u << (f - 3));
		} else if (wt == 2 && f == 2) {
			if (!varint(p, e, v) || (uint64_t)(e - p) < v) fail("bad s_info in " + name);
			const uint8_t* e2 = p + v; uint64_t x;
			while (p < e2) { if (!varint(p, e2, x)) fail("bad s_info in " + name); if (!has) { has = 1; sid = (uint32_t)x; } }
		} else if (!skip(p, e, wt)) fail("bad s_info in " + name);
	}
	if (!seen_first) { seen_first = true; first_index = index; }
	if (has) {
		if (sid == 0 && !seen_ref0) { seen_ref0 = true; ref0_index = index; }
		b.any_sid = true;
		if (b.ssid.size() < b.sflags.size()) b.ssid.resize(b.sflags.size(), 0);
		b.ssid.push_back(sid);
	}
	b.sflags.push_back(flags);
	if (b.keep_index) b.sindex.push_back(index);
}

void parse_vertex(const uint8_t* p, const uint8_t* e, BlockSoA& b, const std::string& name) {
	uint32_t id = 0, off = 0, len = 0, cls = 0; uint8_t has_cls = 0;
	uint32_t first_index = 0, ref0_index = 0; bool seen_first = false, seen_ref0 = false;
	b.sbegin.push_back(b.sflags.size());
	while (p < e) {
		uint64_t tag, v;
		if (!varint(p, e, tag)) fail("bad vertex in " + name);
		unsigned f = (unsigned)(tag >> 3), wt = (unsigned)(tag & 7);
		if (wt == 0 && f >= 1 && f <= 4) {
			if (!varint(p, e, v)) fail("bad vertex in " + name);
			if (f == 1) id = (uint32_t)v; else if (f == 2) off = (uint32_t)v; else if (f == 3) len = (uint32_t)v;
			else if (!has_cls) { has_cls = 1; cls = (uint32_t)v; }
		} else if (wt == 2 && (f == 4 || f == 5)) {
			if (!varint(p, e, v) || (uint64_t)(e - p) < v) fail("bad vertex in " + name);
			const uint8_t* e2 = p + v;
			if (f == 5) parse_sinfo(p, e2, b, name, first_index, ref0_index, seen_first, seen_ref0);
			else { const uint8_t* q = p; uint64_t x; while (q < e2) { if (!varint(q, e2, x)) fail("bad vertex in " + name); if (!has_cls) { has_cls = 1; cls = (uint32_t)x; } } }
			p = e2;
		} else if (!skip(p, e, wt)) fail("bad vertex in " + name);
	}
	b.id.push_back(id); b.offset.push_back(off); b.length.push_back(len); b.cls.push_back(cls); b.has_cls.push_back(has_cls);
	b.first_index.push_back(first_index); b.ref0_index.push_back(ref0_index);
	if (has_cls) b.any_cls = true;
}

void inflate_file(const std::string& path, std::vector<uint8_t>& out) {
	Mapped m(path);
	z_stream zs; memset(&zs, 0, sizeof zs);
	if (inflateInit2(&zs, 15 + 32) != Z_OK) fail("zlib init failed");
	out.resize(std::max<size_t>(m.n * 6, 1 << 16));
	zs.next_in = (Bytef*)m.p; size_t fed = 0; zs.avail_in = 0; size_t produced = 0;
	int ret = Z_OK;
	while (ret != Z_STREAM_END) {
		if (zs.avail_in == 0 && fed < m.n) { size_t chunk = std::min<size_t>(m.n - fed, 1u << 30); zs.next_in = (Bytef*)m.p + fed; zs.avail_in = (uInt)chunk; fed += chunk; }
		if (produced == out.size()) out.resize(out.size() * 2);
		size_t room = std::min<size_t>(out.size() - produced, 1u << 30);
		zs.next_out = out.data() + produced; zs.avail_out = (uInt)room;
		ret = inflate(&zs, Z_NO_FLUSH);
		produced += room - zs.avail_out;
		if (ret != Z_OK && ret != Z_STREAM_END) { inflateEnd(&zs); fail("gzip stream corrupt: " + path); }
		if (ret == Z_OK && zs.avail_in == 0 && fed >= m.n && zs.avail_out != 0) { inflateEnd(&zs); fail("gzip stream truncated: " + path); }
	}
	inflateEnd(&zs);
	out.resize(produced);
}

void load_block(const std::string& path, BlockSoA& b) {
	std::vector<uint8_t> raw; inflate_file(path, raw);
	const uint8_t* p = raw.data(); const uint8_t* e = p + raw.size();
	uint64_t count;
	if (!varint(p, e, count)) fail("bad frame in " + path);
	while (count) {
		for (uint64_t i = 0; i < count; i++) {
			uint64_t len;
			if (!varint(p, e, len) || (uint64_t)(e - p) < len) fail("bad frame in " + path);
			const uint8_t* me = p + len;
			while (p < me) {   // VariantGraphVertexList: repeated vertex = 1
				uint64_t tag, v;
				if (!varint(p, me, tag)) fail("bad list in " + path);
				if ((tag >> 3) == 1 && (tag & 7) == 2) {
					if (!varint(p, me, v) || (uint64_t)(me - p) < v) fail("bad list in " + path);
					parse_vertex(p, p + v, b, path); p += v;
				} else if (!skip(p, me, (unsigned)(tag & 7))) fail("bad list in " + path);
			}
		}
		if (p >= e || !varint(p, e, count)) break;
	}
	b.sbegin.push_back(b.sflags.size());
}

// ------------------------------------------------------------------ CQF image -> (key, value bit, count)
struct CqfEntry { uint64_t key; uint32_t inplace; uint64_t count; };

inline uint64_t unhash40(uint64_t key, uint64_t mask) {   // inverse of the invertible hash (hashutil.c:146-182)
	uint64_t tmp;
	tmp = (key - (key << 31)); key = (key - (tmp << 31)) & mask;
	tmp = key ^ key >> 28; key = key ^ tmp >> 28;
	key = (key * 14933078535860113213ull) & mask;
	tmp = key ^ key >> 14; tmp = key ^ tmp >> 14; tmp = key ^ tmp >> 14; key = key ^ tmp >> 14;
	key = (key * 15244667743933553977ull) & mask;
	tmp = key ^ key >> 24; key = key ^ tmp >> 24;
	tmp = ~key; tmp = ~(key - (tmp << 21)); tmp = ~(key - (tmp << 21)); key = ~(key - (tmp << 21)) & mask;
	return key;
}

void read_cqf(const std::string& path, std::vector<CqfEntry>& out, uint64_t& ndistinct) {
	Mapped m(path);
	if (m.n < 128) fail("truncated " + path);
	auto md = [&](size_t off) { uint64_t v; memcpy(&v, m.p + off, 8); return v; };
	if (md(0) != 1018874902021329732ULL) fail("bad CQF magic in " + path);
	uint32_t hash_mode; memcpy(&hash_mode, m.p + 8, 4);
	if (hash_mode != 1) fail("CQF is not in invertible-hash mode: " + path);
	const uint64_t total = md(16), nslots = md(32), xnslots = md(40), key_bits = md(48), value_bits = md(56),
	               rbits = md(64), bps = md(72), nblocks = md(96);
	ndistinct = md(112);
	const uint64_t stride = 18 + 8 * bps;                  // packed qfblock: u16 offset, u64 occupieds, u64 runends, 64 slots
	if (128 + total > m.n || nblocks > total / stride || bps != rbits + value_bits || bps > 57) fail("inconsistent CQF header in " + path);
	// every slot / occupied / runend bit the scan can touch lies inside the mapped blocks; key = quotient | remainder
	if (bps == 0 || rbits == 0 || value_bits >= bps || nslots == 0 || (nslots & (nslots - 1)) || nslots > xnslots || xnslots > nblocks * 64 ||
	    key_bits > 64 || key_bits != (uint64_t)__builtin_ctzll(nslots) + rbits)
		fail("inconsistent CQF header in " + path);
	const uint8_t* blocks = m.p + 128;
	auto slot = [&](uint64_t i) -> uint64_t {
		if (i >= xnslots) fail("CQF run reads past the last slot in " + path);
		const uint8_t* b = blocks + (i >> 6) * stride + 18; uint64_t bit = (i & 63) * bps;
		uint64_t lo = 0; size_t avail = (size_t)((blocks + total) - (b + bit / 8)); memcpy(&lo, b + bit / 8, std::min<size_t>(8, avail));
		return (lo >> (bit & 7)) & ((1ULL << bps) - 1);
	};
	auto occupied = [&](uint64_t i) { uint64_t w; memcpy(&w, blocks + (i >> 6) * stride + 2, 8); return (w >> (i & 63)) & 1; };
	auto runend = [&](uint64_t i) { uint64_t w; memcpy(&w, blocks + (i >> 6) * stride + 10, 8); return (w >> (i & 63)) & 1; };
	const uint64_t kmask = key_bits >= 64 ? ~0ULL : (1ULL << key_bits) - 1;
	uint64_t next_free = 0;
	for (uint64_t q = 0; q < nslots; q++) {
		if ((q & 63) == 0) { uint64_t w; memcpy(&w, blocks + (q >> 6) * stride + 2, 8); if (!w) { q += 63; continue; } }
		if (!occupied(q)) continue;
		uint64_t s = std::max(q, next_free), e = s;
		while (e < xnslots && !runend(e)) e++;
		if (e >= xnslots) fail("unterminated CQF run in " + path);
		// a run is a sequence of (slot value, counter digits) groups; see encode_counter gqf.c:1043-1108
		for (uint64_t i = s; i <= e;) {
			const uint64_t rem = slot(i);
			uint64_t count, last;
			if (i == e) { count = 1; last = i; }
			else {
				uint64_t d = slot(i + 1);
				if (i + 1 == e) { count = d == rem ? 2 : 1; last = d == rem ? i + 1 : i; }
				else if (rem > 0 && d >= rem) { count = d == rem ? 2 : 1; last = d == rem ? i + 1 : i; }
				else if (rem > 0 && d == 0 && slot(i + 2) == rem) { count = 3; last = i + 2; }
				else if (rem == 0 && d == 0) { if (slot(i + 2) == 0) { count = 3; last = i + 2; } else { count = 2; last = i + 1; } }
				else {
					const uint64_t base = (1ULL << bps) - (rem ? 2 : 1);
					uint64_t acc = 0, j = i + 1;
					while (d != rem && j != e) { if (d > rem) d--; if (d && rem) d--; acc = acc * base + d; j++; d = slot(j); }
					if (rem) { count = acc + 3; last = j; }
					else if (j == e || slot(j + 1) != 0) { count = 1; last = i; }
					else { count = acc + 4; last = j + 1; }
				}
			}
			uint64_t h = (q << rbits) | (rem >> value_bits);
			out.push_back(CqfEntry{unhash40(h, kmask), (uint32_t)(rem & ((1ULL << value_bits) - 1)), count});
			i = last + 1;
		}
		next_free = e + 1;
	}
}

}  // namespace

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
PhaseClock::PhaseClock() : on(getenv("VSGPU_TRACE") != nullptr), t0(now_s()) {}
void PhaseClock::lap(const char* what) { if (!on) return; const double t = now_s(); fprintf(stderr, "[vsgpu trace] open: %-28s %.3f s\n", what, t - t0); t0 = t; }

// Every vertex names a slice of seq_buffer.sdsl; the row materialiser and the render / copy kernels
// read those slices without further checks, so a slice past the end is refused here.
void check_seq_ranges(const SerData& d) {
	const uint64_t n = d.seq.size();
	for (size_t v = 0; v < d.v_offset.size(); v++)
		if ((uint64_t)d.v_offset[v] + d.v_length[v] > n) fail("vertex " + std::to_string(v) + " names sequence [" + std::to_string(d.v_offset[v]) + ", +" + std::to_string(d.v_length[v]) + ") beyond seq_buffer.sdsl (" + std::to_string(n) + " symbols)");
}

void load_ser(const std::string& prefix, SerData& d) {
	PhaseClock pc;
	// ---- sampleid_map.lst
	{
		std::ifstream f(prefix + "/sampleid_map.lst");
		if (!f.good()) fail("cannot open " + prefix + "/sampleid_map.lst");
		std::string name; uint64_t ns = 0; uint32_t id;
		f >> d.chr >> d.ref_length; f >> name >> ns;
		if (ns == 0 || ns > (1u << 30)) fail("bad sample count in sampleid_map.lst");
		d.num_samples = (uint32_t)ns; d.sample_names.assign(ns, std::string());
		uint64_t seen = 0;
		while (f >> name >> id) { if (id >= ns) fail("sample id out of range in sampleid_map.lst"); d.sample_names[id] = name; seen++; }
		if (seen != ns) fail("Num samples is not equal to num entries in samples file.");
	}
	// ---- position index
	{
		std::vector<uint64_t> w;
		d.index_bits = decode_rrr127(prefix + "/index.sdsl", w);
		for (uint64_t i = 0; i < w.size(); i++) for (uint64_t bits = w[i]; bits; bits &= bits - 1) d.index_ones.push_back((uint32_t)(i * 64 + __builtin_ctzll(bits)));
		std::string nm = prefix + "/ref_node_id.sdsl"; Mapped m(nm); Cursor c{m.p, m.p + m.n, nm};
		PackedInts nl = read_iv0(c);
		d.node_list.resize(nl.count);
		for (uint64_t i = 0; i < nl.count; i++) d.node_list[i] = (uint32_t)nl[i];
	}
	pc.lap("sample map + position index");
	// ---- sequence buffer
	{
		std::string nm = prefix + "/seq_buffer.sdsl"; Mapped m(nm); Cursor c{m.p, m.p + m.n, nm};
		PackedInts sb = read_iv0(c);
		d.seq.resize(sb.count);
		for (uint64_t i = 0; i < sb.count; i++) d.seq[i] = (uint8_t)sb[i];
	}
	pc.lap("sequence buffer");
	// ---- sample classes
	{
		std::vector<uint64_t>& w = d.sample_vector;
		struct stat st; std::string nm = prefix + "/sample_vector.sdsl";
		if (stat(nm.c_str(), &st) != 0) fail("cannot open " + nm);
		d.sample_vector_bits = decode_rrr127(nm, w);
	}
	pc.lap("sample classes (rrr)");
	// ---- vertex blocks, decoded in parallel
	{
		std::vector<std::string> files;
		for (uint64_t b = 0;; b++) { std::string nm = prefix + "/vertex_list_" + std::to_string(b) + ".proto"; struct stat st; if (stat(nm.c_str(), &st) != 0) break; files.push_back(nm); }
		if (files.empty()) fail("no vertex_list_<k>.proto under " + prefix);
		std::vector<BlockSoA> blocks(files.size());
		std::atomic<size_t> next{0}; std::string err; std::atomic<bool> bad{false};
		unsigned nt = std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), (unsigned)files.size());
		std::vector<std::thread> th;
		for (unsigned t = 0; t < nt; t++) th.emplace_back([&]() {
			for (size_t i; (i = next++) < files.size();) { try { load_block(files[i], blocks[i]); } catch (const std::exception& e) { if (!bad.exchange(true)) err = e.what(); } }
		});
		for (auto& t : th) t.join();
		if (bad) throw std::runtime_error(err);
		pc.lap("vertex blocks: decode");
		// blocks -> one SoA: offsets first, then every block copies its part in parallel
		const size_t nb = blocks.size();
		std::vector<uint64_t> vbase(nb + 1, 0), sbase(nb + 1, 0);
		for (size_t i = 0; i < nb; i++) { vbase[i + 1] = vbase[i] + blocks[i].id.size(); sbase[i + 1] = sbase[i] + blocks[i].sflags.size(); }
		const uint64_t nv = vbase[nb], ns = sbase[nb];
		if (nv >= 0xFFFFFFFFull) fail("too many vertices");
		d.num_vertices = (uint32_t)nv;
		bool any_sid = false, any_cls = false;
		for (auto& b : blocks) { any_sid |= b.any_sid; any_cls |= b.any_cls; }
		if (any_sid && any_cls) fail("vertex blocks mix sample-class and explicit-id encodings");
		d.class_mode = !any_sid;
		d.v_offset.resize(nv); d.v_length.resize(nv); d.v_class.resize(nv); d.v_sinfo_begin.resize(nv + 1);
		d.v_first_index.resize(nv); d.v_ref0_index.resize(nv);
		d.s_flags.resize(ns);
		if (any_sid) d.s_sample_id.resize(ns);
		std::atomic<size_t> nextb{0}; std::atomic<bool> order_bad{false};
		std::vector<std::thread> th2;
		for (unsigned t = 0; t < nt; t++) th2.emplace_back([&]() {
			for (size_t i; (i = nextb++) < nb;) {
				BlockSoA& b = blocks[i];
				const uint64_t v0 = vbase[i], s0 = sbase[i];
				for (size_t j = 0; j < b.id.size(); j++) {
					if (b.id[j] != v0 + j) order_bad = true;
					d.v_sinfo_begin[v0 + j] = s0 + b.sbegin[j] - b.sbegin[0];
				}
				if (!b.id.empty()) {
					memcpy(&d.v_offset[v0], b.offset.data(), b.id.size() * 4); memcpy(&d.v_length[v0], b.length.data(), b.id.size() * 4);
					memcpy(&d.v_class[v0], b.cls.data(), b.id.size() * 4);
					memcpy(&d.v_first_index[v0], b.first_index.data(), b.id.size() * 4); memcpy(&d.v_ref0_index[v0], b.ref0_index.data(), b.id.size() * 4);
				}
				if (!b.sflags.empty()) {
					memcpy(&d.s_flags[s0], b.sflags.data(), b.sflags.size());
					if (any_sid && !b.ssid.empty()) memcpy(&d.s_sample_id[s0], b.ssid.data(), b.ssid.size() * 4);   // shorter only if its tail had no ids: stays 0
				}
				b = BlockSoA();
			}
		});
		for (auto& t : th2) t.join();
		if (order_bad) fail("vertex ids are not dense/in order in the vertex blocks");
		d.v_sinfo_begin[nv] = ns;
		check_seq_ranges(d);
	}
	pc.lap("vertex blocks: concatenate");
	// ---- topology
	{
		std::vector<CqfEntry> ents;
		read_cqf(prefix + "/adj_list.cqf", ents, d.cqf_distinct);
		pc.lap("cqf scan");
		std::vector<uint32_t> aux, lens;
		auto read_iv32 = [&](const std::string& nm, std::vector<uint32_t>& v) {
			Mapped m(nm); Cursor c{m.p, m.p + m.n, nm};
			uint64_t bits = c.u64(); const uint64_t* w = c.words((bits + 63) / 64);
			v.resize(bits / 32); if (!v.empty()) memcpy(v.data(), w, v.size() * 4);
		};
		read_iv32(prefix + "/aux_vertex_list.sdsl", aux);
		read_iv32(prefix + "/aux_vertex_list_lengths.sdsl", lens);
		std::vector<uint64_t> aux_begin(lens.size() + 1, 0);
		for (size_t i = 0; i < lens.size(); i++) aux_begin[i + 1] = aux_begin[i] + lens[i];
		if (aux_begin.back() > aux.size()) fail("aux_vertex_list shorter than its lengths vector");
		const uint32_t nv = d.num_vertices;
		std::vector<uint32_t> deg(nv, 0);
		for (auto& e : ents) {
			if (e.key >= nv) fail("adjacency key beyond the vertex table");
			if (e.inplace) deg[e.key] = 1;
			else { if (e.count < 1 || e.count > lens.size()) fail("aux pointer out of range in adj_list.cqf"); deg[e.key] = lens[e.count - 1]; }
		}
		d.adj_begin.assign(nv + 1, 0);
		for (uint32_t v = 0; v < nv; v++) d.adj_begin[v + 1] = d.adj_begin[v] + deg[v];
		d.adj.resize(d.adj_begin[nv]);
		auto rows = [&](uint64_t i0, uint64_t i1) {
			std::unordered_set<uint32_t> s;
			for (uint64_t i = i0; i < i1; i++) {
				const CqfEntry& e = ents[i];
				uint32_t* dst = &d.adj[d.adj_begin[e.key]];
				if (e.inplace) { dst[0] = (uint32_t)e.count; continue; }
				// Graph::Graph(prefix) (graph.h:162-171) re-inserts the serialised ids into a fresh
				// std::unordered_set; what the operators then see is that set's iteration order.
				std::unordered_set<uint32_t>().swap(s);               // fresh: the bucket count history is part of the order
				for (uint64_t j = aux_begin[e.count - 1]; j < aux_begin[e.count]; j++) s.insert(aux[j]);
				uint32_t k = 0;
				for (uint32_t n : s) dst[k++] = n;
				for (uint32_t j = k; j < deg[e.key]; j++) dst[j] = UINT32_MAX;   // duplicate ids inside one aux list: shrink the row
			}
		};
		{
			const unsigned nt = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), (unsigned)(ents.size() / 65536 + 1)));
			std::vector<std::thread> th;
			const uint64_t per = (ents.size() + nt - 1) / nt;
			for (unsigned t = 0; t < nt; t++) { const uint64_t a0 = t * per, a1 = std::min<uint64_t>(ents.size(), a0 + per); if (a0 < a1) th.emplace_back(rows, a0, a1); }
			for (auto& t : th) t.join();
		}
		for (uint32_t n : d.adj) if (n != UINT32_MAX && n >= nv) fail("neighbour id beyond the vertex table");
	}
	pc.lap("adjacency rows");
}

// sample_info.index of every s_info, in the order of SerData::v_sinfo_begin: a second pass over the
// vertex blocks, made only by the operators that work in a sample's own coordinates (t3); vsgpu_open
// keeps just the first index of every vertex.
void load_sample_indexes(const std::string& prefix, uint64_t expect, std::vector<uint32_t>& out) {
	std::vector<std::string> files;
	for (uint64_t b = 0;; b++) { std::string nm = prefix + "/vertex_list_" + std::to_string(b) + ".proto"; struct stat st; if (stat(nm.c_str(), &st) != 0) break; files.push_back(nm); }
	std::vector<BlockSoA> blocks(files.size());
	for (auto& b : blocks) b.keep_index = true;
	std::atomic<size_t> next{0}; std::string err; std::atomic<bool> bad{false};
	unsigned nt = std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), (unsigned)std::max<size_t>(files.size(), 1));
	std::vector<std::thread> th;
	for (unsigned t = 0; t < nt; t++) th.emplace_back([&]() {
		for (size_t i; (i = next++) < files.size();) { try { load_block(files[i], blocks[i]); } catch (const std::exception& e) { if (!bad.exchange(true)) err = e.what(); } }
	});
	for (auto& t : th) t.join();
	if (bad) throw std::runtime_error(err);
	uint64_t total = 0;
	for (auto& b : blocks) total += b.sindex.size();
	if (total != expect) fail("vertex blocks changed since the index was opened (" + std::to_string(total) + " sample entries, expected " + std::to_string(expect) + ")");
	out.resize(total);
	uint64_t at = 0;
	for (auto& b : blocks) { if (!b.sindex.empty()) memcpy(&out[at], b.sindex.data(), b.sindex.size() * 4); at += b.sindex.size(); b = BlockSoA(); }
}

}  // namespace vsgpu
