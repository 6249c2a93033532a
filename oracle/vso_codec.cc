// ORACLE — test infrastructure only (see vso.h).  Writers + readers for the ser/ directory:
//   * sdsl-lite containers (NOT vendored by the reference; layouts restated from the published
//     sdsl-lite v2.1 sources from memory — "parity unpinned"):
//       int_vector<0>   u64 size-in-bits, u8 width, ceil(bits/64) u64 words
//       int_vector<32>, bit_vector: u64 size-in-bits, words
//       rrr_vector<127, int_vector<>, 32>: u64 size, bt, btnr, btnrp, rank samples, invert
//   * vertex blocks: gzip( varint64 1 · varint32 len · VariantGraphVertexList )  (stream.hpp:25-52,
//     variantgraphvertex.proto:6-26), proto3 wire format.
#include "vso.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <zlib.h>

namespace vso {
namespace codec {

// ------------------------------------------------------------------ small file helpers
namespace {
struct Out {
	std::string buf;
	void u64(uint64_t v) { buf.append((const char*)&v, 8); }
	void u8(uint8_t v) { buf.push_back((char)v); }
	void words(const std::vector<uint64_t>& w) { if (!w.empty()) buf.append((const char*)w.data(), w.size() * 8); }
	bool save(const std::string& path) {
		FILE* f = fopen(path.c_str(), "wb");
		if (!f) return false;
		bool ok = buf.empty() || fwrite(buf.data(), buf.size(), 1, f) == 1;
		fclose(f);
		return ok;
	}
};
struct In {
	std::vector<uint8_t> buf; size_t p = 0; bool ok = true;
	bool load(const std::string& path) {
		FILE* f = fopen(path.c_str(), "rb");
		if (!f) return false;
		fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
		buf.resize(n);
		bool r = n == 0 || fread(buf.data(), n, 1, f) == 1;
		fclose(f);
		return r;
	}
	uint64_t u64() { uint64_t v = 0; if (p + 8 > buf.size()) { ok = false; return 0; } memcpy(&v, &buf[p], 8); p += 8; return v; }
	uint8_t u8() { if (p + 1 > buf.size()) { ok = false; return 0; } return buf[p++]; }
	void words(std::vector<uint64_t>& w, uint64_t n) {
		if (p + n * 8 > buf.size()) { ok = false; return; }
		w.resize(n); if (n) memcpy(w.data(), &buf[p], n * 8); p += n * 8;
	}
};

// packed fixed-width array <-> words
void pack_bits(const std::vector<uint64_t>& vals, uint8_t width, std::vector<uint64_t>& words) {
	uint64_t bits = (uint64_t)vals.size() * width;
	BitVec bv; bv.resize(bits);
	for (size_t i = 0; i < vals.size(); i++) bv.set_int(i * width, vals[i], width);
	words = bv.w;
}
void unpack_bits(const std::vector<uint64_t>& words, uint64_t nbits, uint8_t width, std::vector<uint64_t>& vals) {
	BitVec bv; bv.nbits = nbits; bv.w = words; bv.w.push_back(0);
	uint64_t n = width ? nbits / width : 0;
	vals.resize(n);
	for (uint64_t i = 0; i < n; i++) vals[i] = bv.get_int(i * width, width);
}
void write_iv0(Out& o, const std::vector<uint64_t>& vals, uint8_t width) {
	std::vector<uint64_t> w; pack_bits(vals, width, w);
	o.u64((uint64_t)vals.size() * width); o.u8(width); o.words(w);
}
bool read_iv0(In& in, std::vector<uint64_t>& vals, uint8_t* width_out = nullptr) {
	uint64_t bits = in.u64(); uint8_t width = in.u8();
	std::vector<uint64_t> w; in.words(w, (bits + 63) / 64);
	if (!in.ok || width == 0 || width > 64) return false;
	unpack_bits(w, bits, width, vals);
	if (width_out) *width_out = width;
	return true;
}
void write_bv(Out& o, const BitVec& bv) { o.u64(bv.nbits); std::vector<uint64_t> w = bv.w; w.resize((bv.nbits + 63) / 64, 0); o.words(w); }
bool read_bv(In& in, BitVec& bv) { bv.nbits = in.u64(); in.words(bv.w, (bv.nbits + 63) / 64); return in.ok; }
inline unsigned hi64(uint64_t x) { return x ? 63 - __builtin_clzll(x) : 0; }   // sdsl::bits::hi
}  // namespace

uint8_t bits_needed(uint64_t maxval) { return (uint8_t)(hi64(maxval) + 1); }

bool write_int_vector0(const std::string& path, const std::vector<uint64_t>& vals, uint8_t width) {
	Out o; write_iv0(o, vals, width); return o.save(path);
}
bool read_int_vector0(const std::string& path, std::vector<uint64_t>& vals, uint8_t* width) {
	In in; if (!in.load(path)) return false; return read_iv0(in, vals, width);
}
bool write_int_vector32(const std::string& path, const std::vector<uint32_t>& vals) {
	Out o; o.u64((uint64_t)vals.size() * 32);
	std::vector<uint64_t> w((vals.size() + 1) / 2, 0);
	if (!vals.empty()) memcpy(w.data(), vals.data(), vals.size() * 4);
	o.words(w);
	return o.save(path);
}
bool read_int_vector32(const std::string& path, std::vector<uint32_t>& vals) {
	In in; if (!in.load(path)) return false;
	uint64_t bits = in.u64(); std::vector<uint64_t> w; in.words(w, (bits + 63) / 64);
	if (!in.ok) return false;
	vals.resize(bits / 32);
	if (!vals.empty()) memcpy(vals.data(), w.data(), vals.size() * 4);
	return true;
}

// ------------------------------------------------------------------ rrr_vector<127>
namespace {
typedef unsigned __int128 u128;
constexpr unsigned BS = 127, SB = 32;   // block size (util.h:28), blocks per superblock (sdsl default t_k)
struct Binom {
	u128 c[BS + 1][BS + 1];
	uint16_t space[BS + 1];
	Binom() {
		for (unsigned n = 0; n <= BS; n++) for (unsigned k = 0; k <= BS; k++) c[n][k] = 0;
		for (unsigned n = 0; n <= BS; n++) { c[n][0] = 1; for (unsigned k = 1; k <= n; k++) c[n][k] = c[n - 1][k - 1] + (k <= n - 1 ? c[n - 1][k] : 0); }
		for (unsigned k = 0; k <= BS; k++) {
			u128 v = c[BS][k];
			if (v == 1) { space[k] = 0; continue; }
			unsigned h = 0; while (v >>= 1) h++;
			space[k] = (uint16_t)(h + 1);
		}
	}
};
const Binom& binom() { static Binom b; return b; }

// combinatorial number system, scanned from bit 0 (rrr_helper::bin_to_nr)
u128 bin_to_nr(u128 bin) {
	if (bin == 0) return 0;
	const Binom& B = binom();
	u128 nr = 0; unsigned k = 0; { u128 t = bin; while (t) { k += (unsigned)(t & 1); t >>= 1; } }
	unsigned nn = BS;
	while (bin != 0) {
		if (bin & 1) { nr += B.c[nn - 1][k]; --k; }
		bin >>= 1; --nn;
	}
	return nr;
}
u128 nr_to_bin(unsigned k, u128 nr) {
	const Binom& B = binom();
	u128 bin = 0;
	for (unsigned p = 0; p < BS && k > 0; p++) {
		unsigned nn = BS - p;
		if (nr >= B.c[nn - 1][k]) { nr -= B.c[nn - 1][k]; --k; bin |= (u128)1 << p; }
	}
	return bin;
}
u128 get_block(const BitVec& bv, uint64_t pos, unsigned len) {
	u128 v = 0;
	if (len > 64) { v = bv.get_int(pos, 64); v |= (u128)bv.get_int(pos + 64, len - 64) << 64; }
	else if (len) v = bv.get_int(pos, len);
	return v;
}
unsigned popc128(u128 v) { return __builtin_popcountll((uint64_t)v) + __builtin_popcountll((uint64_t)(v >> 64)); }
}  // namespace

bool write_rrr127(const std::string& path, const BitVec& bvin) {
	const Binom& B = binom();
	BitVec bv = bvin; bv.w.resize((bv.nbits + 63) / 64 + 2, 0);
	const uint64_t size = bv.nbits;
	const uint64_t nbt = (size + BS) / BS;
	std::vector<uint64_t> bt(nbt, 0);
	uint64_t pos = 0, i = 0, btnr_pos = 0, sum_rank = 0;
	while (pos + BS <= size) { unsigned x = popc128(get_block(bv, pos, BS)); bt[i++] = x; sum_rank += x; btnr_pos += B.space[x]; pos += BS; }
	if (pos < size) { unsigned x = popc128(get_block(bv, pos, (unsigned)(size - pos))); bt[i++] = x; sum_rank += x; btnr_pos += B.space[x]; }
	const uint64_t nsb = (nbt + SB - 1) / SB;
	BitVec btnr; btnr.resize(std::max<uint64_t>(btnr_pos, 64)); btnr.w.resize(btnr.w.size() + 2, 0);
	const uint8_t w_btnrp = (uint8_t)(hi64(btnr_pos) + 1), w_rank = (uint8_t)(hi64(sum_rank) + 1);
	std::vector<uint64_t> btnrp(nsb, 0), rank(nsb + ((size % (SB * BS)) > 0 ? 1 : 0), 0);
	BitVec invert; invert.resize(nsb);
	pos = 0; i = 0; btnr_pos = 0; sum_rank = 0;
	bool inv = false;
	const u128 mask = (((u128)1) << BS) - 1;
	auto put = [&](u128 nr, unsigned sp) {
		if (sp > 64) { btnr.set_int(btnr_pos, (uint64_t)nr, 64); btnr.set_int(btnr_pos + 64, (uint64_t)(nr >> 64), sp - 64); }
		else if (sp) btnr.set_int(btnr_pos, (uint64_t)nr, sp);
	};
	while (pos + BS <= size) {
		if (i % SB == 0) {
			btnrp[i / SB] = btnr_pos; rank[i / SB] = sum_rank;
			if (i + SB <= nbt) {
				unsigned gt_half = 0;
				for (uint64_t j = i; j < i + SB; j++) if (bt[j] > BS / 2) gt_half++;
				if (gt_half > SB / 2) { invert.set(i / SB, 1); for (uint64_t j = i; j < i + SB; j++) bt[j] = BS - bt[j]; inv = true; }
				else inv = false;
			} else inv = false;
		}
		unsigned x = (unsigned)bt[i++];
		unsigned sp = B.space[x];
		sum_rank += inv ? (BS - x) : x;
		if (sp) { u128 bin = get_block(bv, pos, BS); if (inv) bin = (~bin) & mask; put(bin_to_nr(bin), sp); }
		btnr_pos += sp; pos += BS;
	}
	if (pos < size) {
		if (i % SB == 0) { btnrp[i / SB] = btnr_pos; rank[i / SB] = sum_rank; invert.set(i / SB, 0); inv = false; }
		unsigned x = (unsigned)bt[i++];
		unsigned sp = B.space[x];
		sum_rank += inv ? (BS - x) : x;
		if (sp) { u128 bin = get_block(bv, pos, (unsigned)(size - pos)); if (inv) bin = (~bin) & mask; put(bin_to_nr(bin), sp); }
		btnr_pos += sp;
	}
	if (!rank.empty()) rank[rank.size() - 1] = sum_rank;
	btnr.w.resize((btnr.nbits + 63) / 64);
	Out o;
	o.u64(size);
	write_iv0(o, bt, 7);
	write_bv(o, btnr);
	write_iv0(o, btnrp, w_btnrp);
	write_iv0(o, rank, w_rank);
	write_bv(o, invert);
	return o.save(path);
}

bool read_rrr127(const std::string& path, BitVec& out) {
	const Binom& B = binom();
	In in; if (!in.load(path)) return false;
	uint64_t size = in.u64();
	std::vector<uint64_t> bt, btnrp, rank; BitVec btnr, invert;
	if (!read_iv0(in, bt) || !read_bv(in, btnr) || !read_iv0(in, btnrp) || !read_iv0(in, rank) || !read_bv(in, invert)) return false;
	btnr.w.resize(btnr.w.size() + 3, 0);
	out.resize(size); out.w.resize(out.w.size() + 2, 0);
	uint64_t btnr_pos = 0, pos = 0;
	const u128 mask = (((u128)1) << BS) - 1;
	for (uint64_t i = 0; pos < size; i++, pos += BS) {
		if (i >= bt.size()) return false;
		unsigned x = (unsigned)bt[i];
		if (x > BS) return false;
		unsigned sp = B.space[x];
		u128 nr = 0;
		if (sp > 64) { nr = btnr.get_int(btnr_pos, 64); nr |= (u128)btnr.get_int(btnr_pos + 64, sp - 64) << 64; }
		else if (sp) nr = btnr.get_int(btnr_pos, sp);
		btnr_pos += sp;
		u128 bin = (x == BS) ? mask : nr_to_bin(x, nr);
		if (i / SB < invert.nbits && invert.get(i / SB)) bin = (~bin) & mask;
		unsigned len = (unsigned)std::min<uint64_t>(BS, size - pos);
		if (len < BS) bin &= (((u128)1) << len) - 1;
		if (len > 64) { out.set_int(pos, (uint64_t)bin, 64); out.set_int(pos + 64, (uint64_t)(bin >> 64), len - 64); }
		else out.set_int(pos, (uint64_t)bin, len);
	}
	out.w.resize((size + 63) / 64);
	return true;
}

// ------------------------------------------------------------------ protobuf wire format
namespace {
inline void put_varint(std::string& s, uint64_t v) { while (v >= 0x80) { s.push_back((char)(v | 0x80)); v >>= 7; } s.push_back((char)v); }
inline unsigned varint_len(uint64_t v) { unsigned n = 1; while (v >= 0x80) { v >>= 7; n++; } return n; }
inline bool get_varint(const uint8_t*& p, const uint8_t* end, uint64_t& v) {
	v = 0; unsigned shift = 0;
	while (p < end && shift < 64) { uint8_t b = *p++; v |= (uint64_t)(b & 0x7F) << shift; if (!(b & 0x80)) return true; shift += 7; }
	return false;
}
void encode_sinfo(const SampleInfo& s, std::string& o) {
	if (s.index) { o.push_back(0x08); put_varint(o, s.index); }
	if (s.has_sid) { o.push_back(0x12); put_varint(o, varint_len(s.sample_id)); put_varint(o, s.sample_id); }
	if (s.phase) { o.push_back(0x18); o.push_back(1); }
	if (s.gt1) { o.push_back(0x20); o.push_back(1); }
	if (s.gt2) { o.push_back(0x28); o.push_back(1); }
}
void encode_vertex(const Vertex& v, std::string& o, std::string& tmp) {
	if (v.vertex_id) { o.push_back(0x08); put_varint(o, v.vertex_id); }
	if (v.offset) { o.push_back(0x10); put_varint(o, v.offset); }
	if (v.length) { o.push_back(0x18); put_varint(o, v.length); }
	if (v.has_class) { o.push_back(0x22); put_varint(o, varint_len(v.class_id)); put_varint(o, v.class_id); }
	for (const auto& s : v.s_info) { tmp.clear(); encode_sinfo(s, tmp); o.push_back(0x2A); put_varint(o, tmp.size()); o.append(tmp); }
}
bool skip_field(const uint8_t*& p, const uint8_t* end, unsigned wt) {
	uint64_t t;
	switch (wt) {
		case 0: return get_varint(p, end, t);
		case 1: if (end - p < 8) return false; p += 8; return true;
		case 2: if (!get_varint(p, end, t) || (uint64_t)(end - p) < t) return false; p += t; return true;
		case 5: if (end - p < 4) return false; p += 4; return true;
	}
	return false;
}
bool decode_sinfo(const uint8_t* p, const uint8_t* end, SampleInfo& s) {
	while (p < end) {
		uint64_t tag, v;
		if (!get_varint(p, end, tag)) return false;
		unsigned f = (unsigned)(tag >> 3), wt = (unsigned)(tag & 7);
		if (f == 1 && wt == 0) { if (!get_varint(p, end, v)) return false; s.index = (uint32_t)v; }
		else if (f == 2 && wt == 2) {
			if (!get_varint(p, end, v) || (uint64_t)(end - p) < v) return false;
			const uint8_t* e2 = p + v; uint64_t id;
			while (p < e2) { if (!get_varint(p, e2, id)) return false; if (!s.has_sid) { s.has_sid = 1; s.sample_id = (uint32_t)id; } }
		}
		else if (f == 2 && wt == 0) { if (!get_varint(p, end, v)) return false; if (!s.has_sid) { s.has_sid = 1; s.sample_id = (uint32_t)v; } }
		else if (f == 3 && wt == 0) { if (!get_varint(p, end, v)) return false; s.phase = v != 0; }
		else if (f == 4 && wt == 0) { if (!get_varint(p, end, v)) return false; s.gt1 = v != 0; }
		else if (f == 5 && wt == 0) { if (!get_varint(p, end, v)) return false; s.gt2 = v != 0; }
		else if (!skip_field(p, end, wt)) return false;
	}
	return true;
}
bool decode_vertex(const uint8_t* p, const uint8_t* end, Vertex& vx) {
	while (p < end) {
		uint64_t tag, v;
		if (!get_varint(p, end, tag)) return false;
		unsigned f = (unsigned)(tag >> 3), wt = (unsigned)(tag & 7);
		if (f == 1 && wt == 0) { if (!get_varint(p, end, v)) return false; vx.vertex_id = (uint32_t)v; }
		else if (f == 2 && wt == 0) { if (!get_varint(p, end, v)) return false; vx.offset = (uint32_t)v; }
		else if (f == 3 && wt == 0) { if (!get_varint(p, end, v)) return false; vx.length = (uint32_t)v; }
		else if (f == 4 && wt == 2) {
			if (!get_varint(p, end, v) || (uint64_t)(end - p) < v) return false;
			const uint8_t* e2 = p + v; uint64_t c;
			while (p < e2) { if (!get_varint(p, e2, c)) return false; if (!vx.has_class) { vx.has_class = true; vx.class_id = (uint32_t)c; } }
		}
		else if (f == 4 && wt == 0) { if (!get_varint(p, end, v)) return false; if (!vx.has_class) { vx.has_class = true; vx.class_id = (uint32_t)v; } }
		else if (f == 5 && wt == 2) {
			if (!get_varint(p, end, v) || (uint64_t)(end - p) < v) return false;
			SampleInfo s; if (!decode_sinfo(p, p + v, s)) return false;
			vx.s_info.push_back(s); p += v;
		}
		else if (!skip_field(p, end, wt)) return false;
	}
	return true;
}
}  // namespace

void encode_vertex_list(const Vertex* v, size_t n, std::string& out) {
	std::string one, tmp;
	for (size_t i = 0; i < n; i++) {
		one.clear(); encode_vertex(v[i], one, tmp);
		out.push_back(0x0A); put_varint(out, one.size()); out.append(one);
	}
}

bool decode_vertex_list(const uint8_t* p, size_t len, std::vector<Vertex>& out) {
	const uint8_t* end = p + len;
	while (p < end) {
		uint64_t tag, v;
		if (!get_varint(p, end, tag)) return false;
		unsigned f = (unsigned)(tag >> 3), wt = (unsigned)(tag & 7);
		if (f == 1 && wt == 2) {
			if (!get_varint(p, end, v) || (uint64_t)(end - p) < v) return false;
			out.emplace_back();
			if (!decode_vertex(p, p + v, out.back())) return false;
			p += v;
		} else if (!skip_field(p, end, wt)) return false;
	}
	return true;
}

bool write_vertex_block(const std::string& path, const Vertex* v, size_t n, int gzip_level) {
	std::string msg; encode_vertex_list(v, n, msg);
	std::string framed; put_varint(framed, 1); put_varint(framed, msg.size()); framed.append(msg);
	z_stream zs; memset(&zs, 0, sizeof zs);
	if (deflateInit2(&zs, gzip_level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
	FILE* f = fopen(path.c_str(), "wb");
	if (!f) { deflateEnd(&zs); return false; }
	std::vector<uint8_t> obuf(1 << 20);
	zs.next_in = (Bytef*)framed.data(); zs.avail_in = 0;
	size_t fed = 0; int ret = Z_OK; bool ok = true;
	do {
		if (zs.avail_in == 0 && fed < framed.size()) {
			size_t chunk = std::min<size_t>(framed.size() - fed, 1u << 30);
			zs.next_in = (Bytef*)framed.data() + fed; zs.avail_in = (uInt)chunk; fed += chunk;
		}
		zs.next_out = obuf.data(); zs.avail_out = (uInt)obuf.size();
		ret = deflate(&zs, fed >= framed.size() ? Z_FINISH : Z_NO_FLUSH);
		size_t have = obuf.size() - zs.avail_out;
		if (have && fwrite(obuf.data(), have, 1, f) != 1) { ok = false; break; }
	} while (ret != Z_STREAM_END && ret != Z_STREAM_ERROR && ret != Z_BUF_ERROR);
	deflateEnd(&zs); fclose(f);
	return ok && ret == Z_STREAM_END;
}

bool read_vertex_block(const std::string& path, std::vector<Vertex>& out) {
	gzFile f = gzopen(path.c_str(), "rb");
	if (!f) return false;
	gzbuffer(f, 1 << 20);
	std::vector<uint8_t> data; std::vector<uint8_t> buf(1 << 22);
	int n;
	while ((n = gzread(f, buf.data(), (unsigned)buf.size())) > 0) data.insert(data.end(), buf.begin(), buf.begin() + n);
	gzclose(f);
	if (n < 0) return false;
	const uint8_t* p = data.data(); const uint8_t* end = p + data.size();
	uint64_t count;
	if (!get_varint(p, end, count)) return false;
	while (count) {   // stream.hpp:88-107: chunks of `count` length-prefixed messages
		for (uint64_t i = 0; i < count; i++) {
			uint64_t len;
			if (!get_varint(p, end, len) || (uint64_t)(end - p) < len) return false;
			if (len > 0 && !decode_vertex_list(p, len, out)) return false;
			p += len;
		}
		if (p >= end || !get_varint(p, end, count)) break;
	}
	return true;
}

}  // namespace codec
}  // namespace vso
