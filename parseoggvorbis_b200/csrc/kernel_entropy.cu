// Entropy decode of Vorbis audio packets on the device (POV_INPUT_PACKETS).
//
// What the reference does per packet with its bit reader — floor1 flags and coded Y values (src/ParseOggVorbis.hpp:478-518),
// residue classification words and VQ entry numbers in cascade order (hpp:696-760), Huffman decode through the codebook
// trees (hpp:347-360) — is done here by ONE THREAD PER PACKET: the bits of a packet are inherently sequential (the position
// of every codeword depends on all the codewords before it), but a batch holds thousands of independent packets.
// The kernel writes exactly what the host front end would have produced for a POV_INPUT_ENTRIES batch: the Y arena, the
// floor_used mask and the residue payload (n_entries | classifications | entry numbers per submap, layout in
// include/pov_synth.h), so that everything downstream (k_residue_apply, k_warp_synth, the staged kernels) is unchanged.
//
// Huffman decode is ours, not the reference's 1-bit tree walk: a first-level table indexed by the next 9 stream bits
// resolves every codeword of <= 9 bits with one load; longer codewords are found by binary search over the codewords
// sorted by their left-aligned value (tables built once per setup by pov_setup_register from the codeword lengths).
// End-of-packet behaviour is the reference's: bits past the end read as zero and decoding goes on (src/Utils.hpp:389-392).
#include "kernels.h"

namespace pov {

namespace {

struct BitReader {
	const uint32_t* w;      // packet words (the packet starts 4-byte aligned; bytes past its end are zero up to the next word)
	uint32_t nwords, idx;
	uint64_t buf;           // next bits, LSb first
	int cnt;                // valid bits in buf
	__device__ __forceinline__ void init(const uint32_t* p, uint32_t nbytes) {
		w = p; nwords = (nbytes + 3u) >> 2; idx = 0; buf = 0; cnt = 0;
	}
	__device__ __forceinline__ void refill() {      // afterwards cnt > 32
		if(cnt <= 32) {
			const uint32_t x = idx < nwords ? __ldg(w + idx) : 0u;
			++idx;
			buf |= (uint64_t) x << cnt;
			cnt += 32;
		}
	}
	__device__ __forceinline__ uint32_t get(int n) {    // 0 <= n <= 32
		refill();
		const uint32_t v = (uint32_t) buf & (n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1u));
		buf >>= n; cnt -= n;
		return v;
	}
};

__device__ __forceinline__ uint32_t huff_decode(BitReader& br, const uint32_t* __restrict__ arena, const DevHuffBook& hb) {
	br.refill();
	const uint32_t f = __ldg(arena + hb.lut_off + ((uint32_t) br.buf & ((1u << POV_HUFF_LUT_BITS) - 1u)));
	if(f & 63u) {
		br.buf >>= (f & 63u); br.cnt -= (int) (f & 63u);
		return f >> 6;
	}
	// longer codeword: the next 32 stream bits, first bit most significant, against the sorted left-aligned codewords
	const uint32_t msb = __brev((uint32_t) br.buf);
	const uint32_t* code = arena + hb.sorted_off;
	uint32_t lo = 0, hi = hb.n_sorted;
	while(hi - lo > 1) {
		const uint32_t mid = (lo + hi) >> 1;
		if(__ldg(code + mid) <= msb) lo = mid; else hi = mid;
	}
	const uint32_t e = __ldg(code + hb.n_sorted + lo);
	br.buf >>= (e & 63u); br.cnt -= (int) (e & 63u);
	return e >> 6;
}

}  // namespace

// ys_off[p] / ent_off[p]: where packet p's Y lists (uint16 index into b.ys) and entries payload (byte offset into
// ent_out) go: capacity-based, computed on the host from the packet's mode. raw_off[p]: byte offset of the packet in
// b.payload (the caller's pov_packet::spec_off, kept apart because the device copy of the descriptors is patched here:
// floor_used, ys_off, spec_off -> entries payload; the kernel may run again on the same batch).
__global__ void __launch_bounds__(128) k_packet_decode(DevBatchView b, pov_packet* __restrict__ packets, uint16_t* __restrict__ ys_out,
                                                       uint8_t* __restrict__ ent_out, const uint64_t* __restrict__ ys_off,
                                                       const uint64_t* __restrict__ ent_off, const uint64_t* __restrict__ raw_off, uint32_t n_packets) {
	const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if(p >= n_packets) return;
	pov_packet pk = packets[p];
	const DevSetup& su = b.setups[b.streams[pk.stream].setup_id];
	const uint32_t C = su.channels;
	const uint32_t half = su.blocksize[su.mode_blockflag[pk.mode]] / 2;
	const DevMapping& mp = su.mappings[su.mode_mapping[pk.mode]];
	const uint32_t* __restrict__ arena = su.huff;
	const DevHuffBook* __restrict__ hbooks = su.hbooks;
	BitReader br;
	br.init(reinterpret_cast<const uint32_t*>(b.payload + raw_off[p]), pk.packet_bytes);
	// packet type, mode number, window flags: already read by the host (hpp:1142-1153)
	br.get(1);
	br.get((int) su.mode_bits);
	if(su.mode_blockflag[pk.mode]) br.get(2);
	uint32_t status = 0;

	// ---- 4.3.2 floor curve decode: flags + coded Ys (hpp:478-518) ----
	uint16_t* yp = ys_out + ys_off[p];
	uint32_t used = 0;
	for(uint32_t c = 0; c < C; ++c) {
		const uint32_t fi = mp.floor_of_ch[c];
		const DevFloorSyntax& fs = su.fsyntax[fi];
		if(br.get(1) == 0) continue;
		used |= 1u << c;
		*yp++ = (uint16_t) br.get(fs.ybits);
		*yp++ = (uint16_t) br.get(fs.ybits);
		for(uint32_t k = 0; k < fs.n_partitions; ++k) {
			const uint32_t cl = fs.partition_class[k];
			const uint32_t bits = fs.class_subclass_bits[cl], csub = (1u << bits) - 1u;
			uint32_t cval = bits ? huff_decode(br, arena, hbooks[fs.class_masterbook[cl]]) : 0u;
			for(uint32_t i = 0; i < fs.class_dim[cl]; ++i) {
				const int book = fs.class_books[cl][cval & csub];
				cval >>= bits;
				uint32_t y = book >= 0 ? huff_decode(br, arena, hbooks[book]) : 0u;
				if(y > 0xFFFFu) { y = 0xFFFFu; status |= POV_PKT_FLOOR_RANGE; }      // such a curve fails hpp:587 anyway
				*yp++ = (uint16_t) y;
			}
		}
	}
	// ---- 4.3.3 nonzero propagate (hpp:1174-1180) ----
	uint32_t prop = used;
	for(uint32_t k = 0; k < mp.n_couplings; ++k) {
		const uint32_t m = mp.coupling_mag[k], a = mp.coupling_ang[k];
		if(((prop >> m) | (prop >> a)) & 1u) prop |= (1u << m) | (1u << a);
	}
	// ---- 4.3.4 residue decode: classifications + entry numbers (hpp:696-760; walk order of Vorbis I 8.6.2) ----
	const bool ent16 = su.entry_bits == 16;
	uint8_t* out = ent_out + ent_off[p];
	for(uint32_t s = 0; s < mp.n_submaps; ++s) {
		uint32_t chmask = 0, nch = 0;
		for(uint32_t c = 0; c < C; ++c) if(mp.mux[c] == s) { if((prop >> c) & 1u) chmask |= 1u << nch; ++nch; }
		const DevResidue& rs = su.residues[mp.submap_residue[s]];
		const uint32_t vch = rs.type == 2 ? 1u : nch, vlen = rs.type == 2 ? nch * half : half;
		if(rs.type == 2) chmask = 1u;                       // hpp:688: always decoded
		const uint32_t lb = min(rs.begin, vlen), le = min(rs.end, vlen), psize = rs.partition_size;
		const uint32_t parts = (le - lb) / psize;
		uint8_t* cls = out + 4;
		const uint32_t cls_bytes = (vch * parts + 3u) & ~3u;
		for(uint32_t i = 0; i < cls_bytes; i += 4) *reinterpret_cast<uint32_t*>(cls + i) = 0u;
		uint8_t* ent = cls + cls_bytes;
		uint32_t ne = 0;
		if(le != lb) {                                       // hpp:704-705: nothing to read otherwise
			const DevHuffBook& cb = hbooks[su.classbook[mp.submap_residue[s]]];
			const uint32_t cw = cb.dim, ncl = rs.n_class;
			for(uint32_t pass = 0; pass < 8; ++pass) {
				uint32_t pc = 0;
				while(pc < parts) {
					if(pass == 0) {
						for(uint32_t j = 0; j < vch; ++j) {
							if(!((chmask >> j) & 1u)) continue;
							uint32_t t = huff_decode(br, arena, cb);
							for(uint32_t i = cw; i > 0; --i) {
								const uint32_t q = t / ncl;
								if(pc + i - 1 < parts) cls[j * parts + pc + i - 1] = (uint8_t) (t - q * ncl);
								t = q;
							}
						}
					}
					for(uint32_t i = 0; i < cw && pc < parts; ++i, ++pc) {
						for(uint32_t j = 0; j < vch; ++j) {
							if(!((chmask >> j) & 1u)) continue;
							const uint32_t book = rs.books[(uint32_t) cls[j * parts + pc] * 8u + pass];
							if(book == POV_NO_BOOK) continue;
							if(book >= su.n_codebooks) { status |= POV_PKT_VQ_ENTRY; continue; }
							const DevHuffBook& vb = hbooks[book];
							if(vb.lookup_type == 0) status |= POV_PKT_VQ_ENTRY;             // hpp:369-370: not a VQ book
							const uint32_t nvec = psize / vb.dim;
							for(uint32_t k = 0; k < nvec; ++k) {
								const uint32_t e = huff_decode(br, arena, vb);
								if(ent16) reinterpret_cast<uint16_t*>(ent)[ne] = (uint16_t) e;
								else reinterpret_cast<uint32_t*>(ent)[ne] = e;
								++ne;
							}
						}
					}
				}
			}
		}
		*reinterpret_cast<uint32_t*>(out) = ne;
		out = ent + (((uint64_t) ne * (ent16 ? 2u : 4u) + 3u) & ~3ull);
	}
	pk.floor_used = (uint16_t) used;
	pk.ys_off = ys_off[p];
	pk.spec_off = ent_off[p];
	packets[p] = pk;
	if(status) atomicOr(&b.status[p], status);
}

cudaError_t launch_packet_decode(const DevBatchView& b, pov_packet* packets, uint16_t* ys_out, uint8_t* ent_out, const uint64_t* ys_off,
                                 const uint64_t* ent_off, const uint64_t* raw_off, uint32_t n_packets, cudaStream_t st, uint64_t* launches) {
	if(n_packets == 0) return cudaSuccess;
	k_packet_decode<<<(n_packets + 127) / 128, 128, 0, st>>>(b, packets, ys_out, ent_out, ys_off, ent_off, raw_off, n_packets);
	if(launches) ++*launches;
	return cudaGetLastError();
}

}  // namespace pov
