// ORACLE — test infrastructure only (see vso.h).  Topology store restatement:
//   include/graph.h (Graph, GraphIterator) and the slice of the Counting Quotient Filter
//   (src/gqf/gqf.c, gqf_file.c, hashutil.c) that Graph depends on.
#include "vso.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>

namespace vso {

// ------------------------------------------------------------------ BitVec
uint64_t BitVec::get_int(uint64_t pos, unsigned len) const {
	uint64_t wi = pos >> 6; unsigned off = pos & 63;
	uint64_t v = w[wi] >> off;
	if (off + len > 64) v |= w[wi + 1] << (64 - off);
	if (len < 64) v &= ((1ULL << len) - 1);
	return v;
}
void BitVec::set_int(uint64_t pos, uint64_t v, unsigned len) {
	uint64_t wi = pos >> 6; unsigned off = pos & 63;
	uint64_t mask = len < 64 ? ((1ULL << len) - 1) : ~0ULL;
	v &= mask;
	w[wi] = (w[wi] & ~(mask << off)) | (v << off);
	if (off + len > 64) {
		unsigned done = 64 - off;
		uint64_t m2 = mask >> done;
		w[wi + 1] = (w[wi + 1] & ~m2) | (v >> done);
	}
}

// ------------------------------------------------------------------ hashes (hashutil.c)
// Thomas Wang's invertible 64-bit integer hash restricted to `mask` (hashutil.c:132-142).
static inline uint64_t hash_64(uint64_t key, uint64_t mask) {
	key = (~key + (key << 21)) & mask;
	key = key ^ key >> 24;
	key = ((key + (key << 3)) + (key << 8)) & mask;
	key = key ^ key >> 14;
	key = ((key + (key << 2)) + (key << 4)) & mask;
	key = key ^ key >> 28;
	key = (key + (key << 31)) & mask;
	return key;
}
// Inverse (hashutil.c:146-182).
static inline uint64_t hash_64i(uint64_t key, uint64_t mask) {
	uint64_t tmp;
	tmp = (key - (key << 31));
	key = (key - (tmp << 31)) & mask;
	tmp = key ^ key >> 28;
	key = key ^ tmp >> 28;
	key = (key * 14933078535860113213ull) & mask;
	tmp = key ^ key >> 14;
	tmp = key ^ tmp >> 14;
	tmp = key ^ tmp >> 14;
	key = key ^ tmp >> 14;
	key = (key * 15244667743933553977ull) & mask;
	tmp = key ^ key >> 24;
	key = key ^ tmp >> 24;
	tmp = ~key;
	tmp = ~(key - (tmp << 21));
	tmp = ~(key - (tmp << 21));
	key = ~(key - (tmp << 21)) & mask;
	return key;
}

// MurmurHash2 64-bit variant A (Austin Appleby, public domain; hashutil.c:23-64).
uint64_t murmur_hash_64a(const void* key, int len, unsigned int seed) {
	const uint64_t m = 0xc6a4a7935bd1e995ULL;
	const int r = 47;
	uint64_t h = seed ^ (len * m);
	const unsigned char* p = (const unsigned char*)key;
	int nblk = len / 8;
	for (int i = 0; i < nblk; i++) {
		uint64_t k; memcpy(&k, p + 8 * i, 8);
		k *= m; k ^= k >> r; k *= m;
		h ^= k; h *= m;
	}
	const unsigned char* t = p + 8 * nblk;
	int rem = len & 7;
	for (int i = rem - 1; i >= 0; i--) h ^= (uint64_t)t[i] << (8 * i);
	if (rem) h *= m;
	h ^= h >> r; h *= m; h ^= h >> r;
	return h;
}

// ------------------------------------------------------------------ CQF file layout (gqf_int.h:37-101)
namespace {
constexpr uint64_t kMagic = 1018874902021329732ULL;   // gqf_int.h:22
constexpr uint32_t kGqfSeed = 2038074761u;            // gqf_cpp.h:27
constexpr unsigned kKeyBits = 40;                     // graph.h:30
constexpr unsigned kValueBits = 1;
constexpr unsigned kBlockHdr = 18;                    // packed u16 offset + u64 occupieds + u64 runends

#pragma pack(push, 1)
struct QfMeta {            // natural alignment of quotient_filter_metadata on x86-64: 128 bytes
	uint64_t magic; uint32_t hash_mode; uint32_t auto_resize; uint64_t total_size_in_bytes;
	uint32_t seed; uint32_t pad0; uint64_t nslots, xnslots, key_bits, value_bits, key_remainder_bits,
	bits_per_slot; uint64_t range_lo, range_hi; uint64_t nblocks, nelts, ndistinct_elts, noccupied_slots;
};
#pragma pack(pop)
static_assert(sizeof(QfMeta) == 128, "qfmetadata is 128 bytes");

// Number of slots encode_counter (gqf.c:1052-1108) uses for (remainder slot value, count).
unsigned counter_len(uint64_t rem, uint64_t count, unsigned bits_per_slot) {
	if (count == 0) return 0;
	if (count <= 2) return (unsigned)count;
	if (count == 3) return 3;
	uint64_t base = (1ULL << bits_per_slot) - 1;
	unsigned n = 1;   // leading remainder
	uint64_t c = count;
	if (rem == 0) { n++; c -= 4; } else { base--; c -= 3; }
	uint64_t digit = 0;
	do {
		digit = c % base; digit++;
		if (rem && digit >= rem) digit++;
		n++; c /= base;
	} while (c);
	if (rem && digit >= rem) n++;
	n++;               // trailing remainder
	return n;
}
// The slots themselves, in memory order (encode_counter fills them back to front).
void counter_slots(uint64_t rem, uint64_t count, unsigned bits_per_slot, std::vector<uint64_t>& out) {
	std::vector<uint64_t> rev;   // in push order (= reverse memory order)
	if (count == 0) return;
	rev.push_back(rem);
	if (count == 1) { }
	else if (count == 2) { rev.push_back(rem); }
	else if (count == 3 && rem == 0) { rev.push_back(rem); rev.push_back(rem); }
	else if (count == 3) { rev.push_back(0); rev.push_back(rem); }
	else {
		uint64_t base = (1ULL << bits_per_slot) - 1;
		uint64_t c = count;
		if (rem == 0) rev.push_back(rem); else base--;
		if (rem) c -= 3; else c -= 4;
		uint64_t digit;
		do {
			digit = c % base; digit++;
			if (rem && digit >= rem) digit++;
			rev.push_back(digit);
			c /= base;
		} while (c);
		if (rem && digit >= rem) rev.push_back(0);
		rev.push_back(rem);
	}
	for (size_t i = rev.size(); i-- > 0;) out.push_back(rev[i]);
}

struct QfGeom {
	uint64_t nslots, xnslots, nblocks, key_remainder_bits, bits_per_slot, block_bytes, total_bytes;
	explicit QfGeom(unsigned log2_slots) {   // qf_init gqf.c:1632-1652
		nslots = 1ULL << log2_slots;
		xnslots = nslots + (uint64_t)(10 * sqrt((double)nslots));
		nblocks = (xnslots + 63) / 64;
		key_remainder_bits = kKeyBits - log2_slots;
		bits_per_slot = key_remainder_bits + kValueBits;
		block_bytes = kBlockHdr + 64 * bits_per_slot / 8;
		total_bytes = nblocks * block_bytes;
	}
};

// --- port of the CQF as a key->(value,count) dictionary + a canonical-layout file writer.
class PortAdjStore : public AdjStore {
public:
	explicit PortAdjStore(unsigned l2) : log2_slots(l2) {}
	uint64_t query(uint64_t key, uint64_t* value_bit) const override {
		auto it = m.find(key);
		if (it == m.end()) return 0;          // qf_query gqf.c:2092-2093 (value untouched)
		*value_bit = it->second.first;
		return it->second.second;
	}
	int insert(uint64_t key, uint64_t value, uint64_t count) override {
		// qf_insert gqf.c:1911-1926: double while >= 75 % of the slots are in use
		QfGeom g(log2_slots);
		if ((double)noccupied >= g.nslots * 0.75) { log2_slots++; recount(); }
		if (count == 0) return 0;
		auto it = m.find(key);
		if (it != m.end() && it->second.first == value) {   // same (key,value): counts add up
			noccupied -= len_of(key, it->second.first, it->second.second);
			it->second.second += count;
			noccupied += len_of(key, value, it->second.second);
			nelts += count;
			return 0;
		}
		m[key] = std::make_pair(value, count);
		noccupied += len_of(key, value, count);
		nelts += count;
		return 0;
	}
	int remove(uint64_t key, uint64_t value) override {     // qf_delete_key_value gqf.c:2021-2028
		auto it = m.find(key);
		if (it == m.end() || it->second.first != value) return 0;
		int freed = (int)len_of(key, value, it->second.second);
		noccupied -= freed; nelts -= it->second.second;
		m.erase(it);
		return freed;
	}
	uint64_t ndistinct() const override { return m.size(); }
	void enumerate(std::vector<std::array<uint64_t, 3>>& out) const override {
		std::vector<std::pair<uint64_t, uint64_t>> hk;
		for (auto& e : m) hk.push_back({(hash_64(e.first, (1ULL << kKeyBits) - 1) << kValueBits) | e.second.first, e.first});
		std::sort(hk.begin(), hk.end());
		for (auto& h : hk) { auto& e = *m.find(h.second); out.push_back({e.first, e.second.first, e.second.second}); }
	}
	bool serialize(const std::string& path) const override;
	unsigned log2_slots;
	std::unordered_map<uint64_t, std::pair<uint64_t, uint64_t>> m;
	uint64_t noccupied = 0, nelts = 0;
private:
	unsigned len_of(uint64_t key, uint64_t value, uint64_t count) const {
		QfGeom g(log2_slots);
		uint64_t h = (hash_64(key, (1ULL << kKeyBits) - 1) << kValueBits) | value;
		return counter_len(h & ((1ULL << g.bits_per_slot) - 1), count, (unsigned)g.bits_per_slot);
	}
	void recount() { noccupied = 0; for (auto& e : m) noccupied += len_of(e.first, e.second.first, e.second.second); }
};

inline void put_slot(uint8_t* blocks, const QfGeom& g, uint64_t index, uint64_t value) {   // set_slot gqf.c:554-575
	uint8_t* blk = blocks + (index / 64) * g.block_bytes + kBlockHdr;
	uint64_t bitpos = (index % 64) * g.bits_per_slot;
	uint8_t* p = blk + bitpos / 8;
	unsigned shift = bitpos % 8;
	unsigned __int128 t = 0;
	unsigned nbytes = (unsigned)((shift + g.bits_per_slot + 7) / 8);
	memcpy(&t, p, nbytes);
	unsigned __int128 mask = (((unsigned __int128)1 << g.bits_per_slot) - 1) << shift;
	t = (t & ~mask) | (((unsigned __int128)value << shift) & mask);
	memcpy(p, &t, nbytes);
}
inline uint64_t get_slot(const uint8_t* blocks, const QfGeom& g, uint64_t index) {
	const uint8_t* blk = blocks + (index / 64) * g.block_bytes + kBlockHdr;
	uint64_t bitpos = (index % 64) * g.bits_per_slot;
	const uint8_t* p = blk + bitpos / 8;
	unsigned shift = bitpos % 8;
	unsigned __int128 t = 0;
	unsigned nbytes = (unsigned)((shift + g.bits_per_slot + 7) / 8);
	memcpy(&t, p, nbytes);
	return (uint64_t)((t >> shift) & (((unsigned __int128)1 << g.bits_per_slot) - 1));
}
inline void set_meta_bit(uint8_t* blocks, const QfGeom& g, uint64_t index, bool runend) {
	uint8_t* blk = blocks + (index / 64) * g.block_bytes;
	uint64_t wv; memcpy(&wv, blk + 2 + (runend ? 8 : 0), 8);
	wv |= 1ULL << (index % 64);
	memcpy(blk + 2 + (runend ? 8 : 0), &wv, 8);
}
inline bool get_meta_bit(const uint8_t* blocks, const QfGeom& g, uint64_t index, bool runend) {
	const uint8_t* blk = blocks + (index / 64) * g.block_bytes;
	uint64_t wv; memcpy(&wv, blk + 2 + (runend ? 8 : 0), 8);
	return (wv >> (index % 64)) & 1;
}

// Canonical quotient-filter layout: runs in quotient order, each at max(quotient, end of the
// previous run + 1), elements of a run in ascending slot value, counters per encode_counter;
// block offset = number of leading slots of the block owned by runs of earlier quotients
// (block_offset gqf.c:577-588).
bool PortAdjStore::serialize(const std::string& path) const {
	QfGeom g(log2_slots);
	std::vector<uint8_t> blocks(g.total_bytes, 0);
	struct E { uint64_t hash, count; };
	std::vector<E> es; es.reserve(m.size());
	for (auto& e : m) es.push_back({(hash_64(e.first, (1ULL << kKeyBits) - 1) << kValueBits) | e.second.first, e.second.second});
	std::sort(es.begin(), es.end(), [](const E& a, const E& b) { return a.hash < b.hash; });
	uint64_t cursor = 0, used = 0, nelts_sum = 0;
	uint64_t next_block = 0;     // next block whose offset field is still to be written
	std::vector<uint64_t> slots;
	size_t i = 0;
	auto flush_offsets_upto = [&](uint64_t q) {
		// every block b with 64*b <= q has all quotients < 64*b placed; offset = max(0, cursor - 64*b)
		while (next_block < g.nblocks && next_block * 64 <= q) {
			uint64_t b0 = next_block * 64;
			uint64_t off = cursor > b0 ? cursor - b0 : 0;
			uint16_t o16 = (uint16_t)std::min<uint64_t>(off, 0xFFFF);
			memcpy(&blocks[next_block * g.block_bytes], &o16, 2);
			next_block++;
		}
	};
	while (i < es.size()) {
		uint64_t q = es[i].hash >> g.bits_per_slot;
		flush_offsets_upto(q);
		slots.clear();
		while (i < es.size() && (es[i].hash >> g.bits_per_slot) == q) {
			counter_slots(es[i].hash & ((1ULL << g.bits_per_slot) - 1), es[i].count, (unsigned)g.bits_per_slot, slots);
			nelts_sum += es[i].count;
			i++;
		}
		uint64_t start = std::max(q, cursor);
		if (start + slots.size() > g.xnslots) { fprintf(stderr, "vso: CQF overflow\n"); return false; }
		for (size_t k = 0; k < slots.size(); k++) put_slot(blocks.data(), g, start + k, slots[k]);
		set_meta_bit(blocks.data(), g, q, false);
		set_meta_bit(blocks.data(), g, start + slots.size() - 1, true);
		cursor = start + slots.size();
		used += slots.size();
	}
	flush_offsets_upto(UINT64_MAX);
	QfMeta md; memset(&md, 0, sizeof md);
	md.magic = kMagic; md.hash_mode = 1; md.auto_resize = 1; md.total_size_in_bytes = g.total_bytes;
	md.seed = kGqfSeed; md.nslots = g.nslots; md.xnslots = g.xnslots; md.key_bits = kKeyBits; md.value_bits = kValueBits;
	md.key_remainder_bits = g.key_remainder_bits; md.bits_per_slot = g.bits_per_slot;
	unsigned __int128 range = (unsigned __int128)g.nslots << g.key_remainder_bits;
	md.range_lo = (uint64_t)range; md.range_hi = (uint64_t)(range >> 64);
	md.nblocks = g.nblocks; md.nelts = nelts_sum; md.ndistinct_elts = m.size(); md.noccupied_slots = used;
	FILE* f = fopen(path.c_str(), "wb");
	if (!f) return false;
	bool ok = fwrite(&md, sizeof md, 1, f) == 1 && fwrite(blocks.data(), blocks.size(), 1, f) == 1;
	fclose(f);
	return ok;
}

// decode_counter gqf.c:1112-1182 over a run [index, run_last]
uint64_t decode_counter_at(const uint8_t* blocks, const QfGeom& g, uint64_t index, uint64_t* remainder, uint64_t* count) {
	auto slot = [&](uint64_t i) { return get_slot(blocks, g, i); };
	auto runend = [&](uint64_t i) { return get_meta_bit(blocks, g, i, true); };
	uint64_t rem = slot(index); *remainder = rem;
	if (runend(index)) { *count = 1; return index; }
	uint64_t digit = slot(index + 1);
	if (runend(index + 1)) { *count = digit == rem ? 2 : 1; return index + (digit == rem ? 1 : 0); }
	if (rem > 0 && digit >= rem) { *count = digit == rem ? 2 : 1; return index + (digit == rem ? 1 : 0); }
	if (rem > 0 && digit == 0 && slot(index + 2) == rem) { *count = 3; return index + 2; }
	if (rem == 0 && digit == 0) {
		if (slot(index + 2) == 0) { *count = 3; return index + 2; }
		*count = 2; return index + 1;
	}
	uint64_t cnt = 0, base = (1ULL << g.bits_per_slot) - (rem ? 2 : 1), end = index + 1;
	while (digit != rem && !runend(end)) {
		if (digit > rem) digit--;
		if (digit && rem) digit--;
		cnt = cnt * base + digit;
		end++; digit = slot(end);
	}
	if (rem) { *count = cnt + 3; return end; }
	if (runend(end) || slot(end + 1) != 0) { *count = 1; return index; }
	*count = cnt + 4; return end + 1;
}
}  // namespace

std::unique_ptr<AdjStore> make_port_adjstore(unsigned log2_slots) {
	return std::unique_ptr<AdjStore>(new PortAdjStore(log2_slots));
}

std::unique_ptr<AdjStore> load_port_adjstore(const std::string& path) {
	FILE* f = fopen(path.c_str(), "rb");
	if (!f) return nullptr;
	QfMeta md;
	if (fread(&md, sizeof md, 1, f) != 1 || md.magic != kMagic) { fclose(f); return nullptr; }
	unsigned l2 = 0; while ((1ULL << l2) < md.nslots) l2++;
	QfGeom g(l2);
	if (g.total_bytes != md.total_size_in_bytes || g.bits_per_slot != md.bits_per_slot || md.hash_mode != 1) { fclose(f); return nullptr; }
	std::vector<uint8_t> blocks(g.total_bytes);
	if (fread(blocks.data(), blocks.size(), 1, f) != 1) { fclose(f); return nullptr; }
	fclose(f);
	auto st = new PortAdjStore(l2);
	// qfi iteration (gqf.c:2207-2436) restated as a linear sweep: occupied quotients in order, each
	// run starts at max(q, previous run end + 1) and ends at the next runend bit.
	uint64_t cursor = 0;
	for (uint64_t q = 0; q < g.nslots; q++) {
		if ((q & 63) == 0) {   // skip empty blocks quickly
			uint64_t occ; memcpy(&occ, &blocks[(q / 64) * g.block_bytes + 2], 8);
			if (occ == 0) { q += 63; continue; }
		}
		if (!get_meta_bit(blocks.data(), g, q, false)) continue;
		uint64_t start = std::max(q, cursor), end = start;
		while (!get_meta_bit(blocks.data(), g, end, true)) end++;
		uint64_t i = start;
		while (i <= end) {
			uint64_t rem, cnt;
			uint64_t last = decode_counter_at(blocks.data(), g, i, &rem, &cnt);
			uint64_t h = (q << g.key_remainder_bits) | (rem >> kValueBits);
			uint64_t key = hash_64i(h, (1ULL << kKeyBits) - 1);
			st->m[key] = std::make_pair(rem & 1, cnt);
			st->nelts += cnt;
			i = last + 1;
		}
		st->noccupied += end - start + 1;
		cursor = end + 1;
	}
	return std::unique_ptr<AdjStore>(st);
}

// ------------------------------------------------------------------ real gqf via dlopen (oracle/_ref)
namespace {
struct RefApi {
	void* h = nullptr;
	bool (*qf_malloc)(void*, uint64_t, uint64_t, uint64_t, int, uint32_t);
	void (*qf_set_auto_resize)(void*, bool);
	int (*qf_insert)(void*, uint64_t, uint64_t, uint64_t, uint8_t);
	uint64_t (*qf_query)(const void*, uint64_t, uint64_t*, uint8_t);
	int (*qf_delete_key_value)(void*, uint64_t, uint64_t, uint8_t);
	uint64_t (*qf_serialize)(const void*, const char*);
	uint64_t (*qf_deserialize)(void*, const char*);
	int64_t (*qf_iterator_from_position)(const void*, void*, uint64_t);
	int (*qfi_get_key)(const void*, uint64_t*, uint64_t*, uint64_t*);
	int (*qfi_next)(void*);
	bool (*qfi_end)(const void*);
	bool (*qf_free)(void*);
};
RefApi* ref_api() {
	static RefApi api; static bool tried = false;
	if (tried) return api.h ? &api : nullptr;
	tried = true;
	std::string path;
	if (const char* e = getenv("VSO_GQF_REF")) path = e;
	else {
		Dl_info info;
		if (dladdr((void*)&ref_api, &info) && info.dli_fname) {
			std::string p = info.dli_fname; size_t s = p.find_last_of('/');
			path = (s == std::string::npos ? std::string(".") : p.substr(0, s)) + "/_ref/libgqf_ref.so";
		}
	}
	void* h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
	if (!h) return nullptr;
#define L(n) *(void**)(&api.n) = dlsym(h, #n); if (!api.n) { dlclose(h); return nullptr; }
	L(qf_malloc) L(qf_set_auto_resize) L(qf_insert) L(qf_query) L(qf_delete_key_value) L(qf_serialize)
	L(qf_deserialize) L(qf_iterator_from_position) L(qfi_get_key) L(qfi_next) L(qfi_end) L(qf_free)
#undef L
	api.h = h;
	return &api;
}
constexpr uint8_t QF_NO_LOCK = 0x01;
class RefAdjStore : public AdjStore {
public:
	RefApi* a; void* qf[3] = {nullptr, nullptr, nullptr};   // struct quotient_filter = 3 pointers
	explicit RefAdjStore(RefApi* api) : a(api) {}
	~RefAdjStore() override { if (qf[1]) a->qf_free(qf); }
	uint64_t query(uint64_t key, uint64_t* v) const override { return a->qf_query(qf, key, v, QF_NO_LOCK); }
	int insert(uint64_t key, uint64_t value, uint64_t count) override { return a->qf_insert(qf, key, value, count, QF_NO_LOCK); }
	int remove(uint64_t key, uint64_t value) override { return a->qf_delete_key_value(qf, key, value, QF_NO_LOCK); }
	uint64_t ndistinct() const override { uint64_t v; memcpy(&v, (const char*)qf[1] + 112, 8); return v; }
	bool serialize(const std::string& path) const override { return a->qf_serialize(qf, path.c_str()) > 0; }
	void enumerate(std::vector<std::array<uint64_t, 3>>& out) const override {
		alignas(16) unsigned char it[128];
		memset(it, 0, sizeof it);
		if (a->qf_iterator_from_position(qf, it, 0) < 0) return;
		while (!a->qfi_end(it)) {
			uint64_t k, v, c;
			a->qfi_get_key(it, &k, &v, &c);
			out.push_back({k, v, c});
			a->qfi_next(it);
		}
	}
};
}  // namespace

std::unique_ptr<AdjStore> make_ref_adjstore(unsigned log2_slots) {
	RefApi* a = ref_api();
	if (!a) return nullptr;
	auto s = new RefAdjStore(a);
	if (!a->qf_malloc(s->qf, 1ULL << log2_slots, kKeyBits, kValueBits, 1 /*QF_HASH_INVERTIBLE*/, kGqfSeed)) { delete s; return nullptr; }
	a->qf_set_auto_resize(s->qf, true);
	return std::unique_ptr<AdjStore>(s);
}
std::unique_ptr<AdjStore> load_ref_adjstore(const std::string& path) {
	RefApi* a = ref_api();
	if (!a) return nullptr;
	auto s = new RefAdjStore(a);
	if (a->qf_deserialize(s->qf, path.c_str()) == 0) { delete s; return nullptr; }
	return std::unique_ptr<AdjStore>(s);
}

// ------------------------------------------------------------------ Graph (graph.h)
Graph::Graph(unsigned log2_slots, bool use_ref_gqf) {
	if (use_ref_gqf) {
		adj = make_ref_adjstore(log2_slots);
		if (!adj) { fprintf(stderr, "vso: reference gqf library (oracle/_ref/libgqf_ref.so) not available\n"); abort(); }
	} else adj = make_port_adjstore(log2_slots);
}

Graph::Graph(const std::string& prefix, bool use_ref_gqf) {
	adj = use_ref_gqf ? load_ref_adjstore(prefix + "/adj_list.cqf") : load_port_adjstore(prefix + "/adj_list.cqf");
	if (!adj) { fprintf(stderr, "vso: can't read %s/adj_list.cqf\n", prefix.c_str()); abort(); }
	std::vector<uint32_t> vertex_list, list_lengths;
	if (!codec::read_int_vector32(prefix + "/aux_vertex_list.sdsl", vertex_list) ||
	    !codec::read_int_vector32(prefix + "/aux_vertex_list_lengths.sdsl", list_lengths)) {
		fprintf(stderr, "vso: can't read aux vertex lists under %s\n", prefix.c_str()); abort();
	}
	uint64_t v_idx = 0;
	for (uint32_t size : list_lengths) {           // graph.h:162-171: re-insert in file order
		vertex_set v_set;
		for (uint64_t pos = v_idx; pos < v_idx + size; ++pos) v_set.insert(vertex_list[pos]);
		aux_vertex_list.emplace_back(v_set);
		v_idx += size;
	}
}

void Graph::serialize(const std::string& prefix) const {
	adj->serialize(prefix + "/adj_list.cqf");
	std::vector<uint32_t> vertex_list, list_lengths;
	for (const auto& list : aux_vertex_list) {     // graph.h:196-200: iteration order of each set
		for (const auto v : list) vertex_list.push_back(v);
		list_lengths.push_back((uint32_t)list.size());
	}
	codec::write_int_vector32(prefix + "/aux_vertex_list.sdsl", vertex_list);
	codec::write_int_vector32(prefix + "/aux_vertex_list_lengths.sdsl", list_lengths);
}

int Graph::add_edge(const vertex s, const vertex d) {
	uint64_t is_inplace = 0;
	vertex val = (vertex)adj->query(s, &is_inplace);
	if (d == 0) return 0;                                       // graph.h:214-215
	if (is_inplace == 1 && val == d) return 0;
	else if (val == 0) { num_edges++; return adj->insert(s, 1, d); }
	else {
		if (is_inplace == 1) {                                    // second neighbour: spill to an aux set
			vertex_set neighbors;
			neighbors.insert(val);
			neighbors.insert(d);
			aux_vertex_list.emplace_back(neighbors);
			uint32_t pointer = (uint32_t)aux_vertex_list.size();
			num_edges++;
			if (adj->remove(s, 1)) return adj->insert(s, 0, pointer);   // replace_key gqf_cpp.h:223-229
			return -1;
		} else {
			if (aux_vertex_list[val - 1].insert(d).second) num_edges++;
		}
	}
	return 0;
}

int Graph::remove_edge(const vertex s, const vertex d) {
	uint64_t is_inplace = 0;
	vertex val = (vertex)adj->query(s, &is_inplace);
	if (val == 0) return 0;
	if (is_inplace == 1) return adj->remove(s, 1);
	for (auto const vertex : aux_vertex_list[val - 1]) {
		if (vertex == d) {                                         // graph.h:253-255 erases begin(), not d
			aux_vertex_list[val - 1].erase(aux_vertex_list[val - 1].begin());
			break;
		}
	}
	return 0;
}

Graph::vertex_set Graph::out_neighbors(const vertex v) const {
	vertex_set neighbor_set;
	uint64_t is_inplace = 0;
	vertex val = (vertex)adj->query(v, &is_inplace);
	if (val == 0) return neighbor_set;
	if (is_inplace == 1) neighbor_set.insert(val);
	else neighbor_set = aux_vertex_list[val - 1];
	return neighbor_set;
}

uint32_t Graph::out_degree(const vertex v) const {
	uint64_t is_inplace = 0;
	vertex val = (vertex)adj->query(v, &is_inplace);
	if (val == 0) return 0;
	if (is_inplace == 1) return 1;
	return (uint32_t)aux_vertex_list[val - 1].size();
}

bool Graph::is_edge(vertex s, vertex d) const {
	uint64_t is_inplace = 0;
	vertex val = (vertex)adj->query(s, &is_inplace);
	if (val == 0) return false;
	if (is_inplace == 1) return val == d;
	return aux_vertex_list[val - 1].find(d) != aux_vertex_list[val - 1].end();
}

Graph::GraphIterator::GraphIterator(const Graph* graph, vertex v, uint64_t radius) {
	g = graph; cur = v; visited.insert(v); r = radius; is_done = false;
	if (radius > 0)
		for (const auto n : g->out_neighbors(v)) q.push_back(std::make_pair(n, (uint64_t)1));
}

void Graph::GraphIterator::operator++() {
	vertex cur_vertex = 0; uint64_t hop = 0;
	while (qh < q.size()) {
		cur_vertex = q[qh].first; hop = q[qh].second;
		if (visited.find(cur_vertex) == visited.end()) { visited.insert(cur_vertex); break; }
		else qh++;
	}
	if (qh >= q.size()) { is_done = true; return; }
	cur = cur_vertex;
	qh++;
	if (hop < r) {
		// graph.h:433-451: neighbours that share an out-neighbour with `cur` go first.  The
		// reference's std::copy(set1.begin(), set2.end(), ...) walks set1 to its null end on
		// libstdc++, i.e. copies all of set1.
		std::vector<vertex> ordered_neighbors;
		auto set1 = g->out_neighbors(cur);
		std::vector<vertex> vec1(set1.begin(), set1.end());
		std::sort(vec1.begin(), vec1.end());
		for (const auto v : set1) {
			auto set2 = g->out_neighbors(v);
			std::vector<vertex> vec2(set2.begin(), set2.end()), intersect;
			std::sort(vec2.begin(), vec2.end());
			std::set_intersection(vec1.begin(), vec1.end(), vec2.begin(), vec2.end(), std::back_inserter(intersect));
			if (intersect.size() > 0) ordered_neighbors.emplace(ordered_neighbors.begin(), v);
			else ordered_neighbors.emplace(ordered_neighbors.end(), v);
		}
		for (const auto v : ordered_neighbors) q.push_back(std::make_pair(v, hop + 1));
		if (qh > (1u << 20) && qh * 2 > q.size()) { q.erase(q.begin(), q.begin() + qh); qh = 0; }
	}
}

}  // namespace vso
