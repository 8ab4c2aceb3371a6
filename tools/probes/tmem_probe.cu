// TMEM probe (sm_100a): (1) thread<->(lane,column) maps of the tcgen05.ld/st shapes, measured by storing lane*1000+col with
// 32x32b and reading back with every other shape (and the reverse); (2) throughput of tcgen05.ld next to ld.shared, alone
// and mixed, with 20 warps per SM as in k_warp_synth.   nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_probe tmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

#define LD_ASM1(shape, num, r, addr) asm volatile("tcgen05.ld.sync.aligned." shape "." num ".b32 {%0}, [%1];" : "=r"(r[0]) : "r"(addr))
#define LD_ASM2(shape, num, r, addr) asm volatile("tcgen05.ld.sync.aligned." shape "." num ".b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr))
#define LD_ASM4(shape, num, r, addr) asm volatile("tcgen05.ld.sync.aligned." shape "." num ".b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr))
#define LD_ASM8(shape, num, r, addr) asm volatile("tcgen05.ld.sync.aligned." shape "." num ".b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr))
#define ST_ASM4(shape, num, r, addr) asm volatile("tcgen05.st.sync.aligned." shape "." num ".b32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory")
#define ST_ASM8(shape, num, r, addr) asm volatile("tcgen05.st.sync.aligned." shape "." num ".b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory")
#define ST_ASM2(shape, num, r, addr) asm volatile("tcgen05.st.sync.aligned." shape "." num ".b32 [%0], {%1,%2};" :: "r"(addr), "r"(r[0]), "r"(r[1]) : "memory")
#define ST_ASM1(shape, num, r, addr) asm volatile("tcgen05.st.sync.aligned." shape "." num ".b32 [%0], {%1};" :: "r"(addr), "r"(r[0]) : "memory")
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// out[test][warp][thread][8]
__global__ void __launch_bounds__(128, 1) k_shapes(uint32_t* out) {
	__shared__ uint32_t s_base;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if(warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&s_base)) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t base = s_base;
	const uint32_t wbase = base + ((uint32_t) (warp * 32) << 16);
	// fill 32 columns: value = (absolute lane) * 1000 + column
	{
		uint32_t v[8];
		for(int c0 = 0; c0 < 32; c0 += 8) {
			for(int i = 0; i < 8; ++i) v[i] = (uint32_t) ((warp * 32 + lane) * 1000 + c0 + i);
			ST_ASM8("32x32b", "x8", v, wbase + c0);
		}
		wait_st();
	}
	__syncwarp();
	uint32_t r[8];
	auto dump = [&](int test, int nreg) {
		wait_ld();
		for(int i = 0; i < 8; ++i) out[((test * 4 + warp) * 32 + lane) * 8 + i] = i < nreg ? r[i] : 0xFFFFFFFFu;
	};
	LD_ASM4("32x32b", "x4", r, wbase); dump(0, 4);
	LD_ASM4("16x64b", "x4", r, wbase); dump(1, 4);                           // 16 lanes x 8 columns
	LD_ASM4("16x64b", "x4", r, wbase + (16u << 16)); dump(2, 4);             // upper 16 lanes
	LD_ASM4("16x128b", "x2", r, wbase); dump(3, 4);                          // 16 lanes x 8 columns
	LD_ASM4("16x256b", "x1", r, wbase); dump(4, 4);                          // 16 lanes x 8 columns
	LD_ASM8("16x256b", "x2", r, wbase); dump(5, 8);                          // 16 lanes x 16 columns
	LD_ASM8("16x128b", "x4", r, wbase + (16u << 16)); dump(6, 8);            // upper 16 lanes x 16 columns
	asm volatile("tcgen05.ld.sync.aligned.16x32bx2.x4.b32 {%0,%1,%2,%3}, [%4], 16;" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(wbase)); dump(7, 4);
	// reverse direction: store with 16x256b.x2 (thread t register i = t*10+i tagged), read with 32x32b
	__syncwarp();
	{
		uint32_t v[8];
		for(int i = 0; i < 8; ++i) v[i] = (uint32_t) (lane * 10 + i + 100000 * (warp + 1));
		ST_ASM8("16x256b", "x2", v, wbase + 32); wait_st();
		for(int i = 0; i < 8; ++i) v[i] += 50000u;
		ST_ASM8("16x256b", "x2", v, wbase + 32 + (16u << 16)); wait_st();
		__syncwarp();
		LD_ASM8("32x32b", "x8", r, wbase + 32); dump(8, 8);
		LD_ASM8("32x32b", "x8", r, wbase + 40); dump(9, 8);
	}
	{
		uint32_t v[8];
		for(int i = 0; i < 8; ++i) v[i] = (uint32_t) (lane * 10 + i + 100000 * (warp + 1));
		ST_ASM8("16x128b", "x4", v, wbase + 48); wait_st();
		for(int i = 0; i < 8; ++i) v[i] += 50000u;
		ST_ASM8("16x128b", "x4", v, wbase + 48 + (16u << 16)); wait_st();
		__syncwarp();
		LD_ASM8("32x32b", "x8", r, wbase + 48); dump(10, 8);
		LD_ASM8("32x32b", "x8", r, wbase + 56); dump(11, 8);
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if(warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(base) : "memory");
}

// throughput: mode bit 0 = TMEM loads, bit 1 = shared loads; `warps` warps per CTA, one CTA per SM
template <int MODE>
__global__ void __launch_bounds__(640, 1) k_bw(uint32_t* out, int iters, long long* cycles) {
	extern __shared__ __align__(16) uint32_t sm[];
	__shared__ uint32_t s_base;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for(int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
	if(warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_base)) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t wbase = s_base + ((uint32_t) ((warp & 3) * 32) << 16);
	{
		uint32_t v[8];
		for(int i = 0; i < 8; ++i) v[i] = lane + i;
		if(warp < 4) for(int c = 0; c < 512; c += 8) ST_ASM8("32x32b", "x8", v, wbase + c);
		wait_st();
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	uint32_t acc = 0;
	const uint4* s4 = reinterpret_cast<const uint4*>(sm) + lane;
	const long long t0 = clock64();
	for(int it = 0; it < iters; ++it) {
		uint32_t r[4], q[4], r2[4], q2[4];
		const uint32_t col = (uint32_t) ((it * 16 + warp * 64) & 511 & ~15);
		if(MODE & 1) { LD_ASM4("32x32b", "x4", r, wbase + col); LD_ASM4("32x32b", "x4", q, wbase + col + 4); LD_ASM4("32x32b", "x4", r2, wbase + col + 8); LD_ASM4("32x32b", "x4", q2, wbase + col + 12); }
		uint4 a = make_uint4(0, 0, 0, 0), b = a, c = a, d = a;
		if(MODE & 2) { a = s4[(it * 32) & 1023]; b = s4[(it * 32 + 256) & 1023]; c = s4[(it * 32 + 512) & 1023]; d = s4[(it * 32 + 768) & 1023]; }
		if(MODE & 1) { wait_ld(); acc += r[0] ^ r[1] ^ r[2] ^ r[3] ^ q[0] ^ q[1] ^ q[2] ^ q[3] ^ r2[0] ^ r2[1] ^ r2[2] ^ r2[3] ^ q2[0] ^ q2[1] ^ q2[2] ^ q2[3]; }
		if(MODE & 2) acc += a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w ^ d.x ^ d.y ^ d.z ^ d.w;
	}
	const long long t1 = clock64();
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
	if(threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if(warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(s_base) : "memory");
}

// exchange trip as in k_warp_synth: STTM.x32 -> (wait::st) -> 2 x LDTM.16dp256bit.x4 -> wait::ld, 20 warps, with FILL dependent FFMAs per trip
template <int FENCE, int FILL>
__global__ void __launch_bounds__(640, 1) k_trip(uint32_t* out, int iters, long long* cycles) {
	__shared__ uint32_t s_base;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if(warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_base)) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t xa = s_base + ((uint32_t) ((warp & 3) * 32) << 16) + 32u * (uint32_t) (warp >> 2);
	float v[32];
	for(int i = 0; i < 32; ++i) v[i] = (float) (lane * 32 + i);
	const long long t0 = clock64();
	for(int it = 0; it < iters; ++it) {
		asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
		             :: "r"(xa), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]),
		                "f"(v[16]), "f"(v[17]), "f"(v[18]), "f"(v[19]), "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]), "f"(v[30]), "f"(v[31]) : "memory");
		if(FENCE) wait_st();
		__syncwarp();
		asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
		             : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]) : "r"(xa));
		asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
		             : "=f"(v[16]), "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31]) : "r"(xa + (16u << 16)));
		asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]), "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15]),
		             "+f"(v[16]), "+f"(v[17]), "+f"(v[18]), "+f"(v[19]), "+f"(v[20]), "+f"(v[21]), "+f"(v[22]), "+f"(v[23]), "+f"(v[24]), "+f"(v[25]), "+f"(v[26]), "+f"(v[27]), "+f"(v[28]), "+f"(v[29]), "+f"(v[30]), "+f"(v[31]));
#pragma unroll
		for(int f = 0; f < FILL; ++f)
#pragma unroll
			for(int i = 0; i < 32; ++i) v[i] = fmaf(v[i], 1.0000001f, 0.5f);
	}
	const long long t1 = clock64();
	float acc = 0; for(int i = 0; i < 32; ++i) acc += v[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = __float_as_uint(acc);
	if(threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if(warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(s_base) : "memory");
}
template <int FENCE, int FILL> static void run_trip(uint32_t* o2, long long* cyc, int warps) {
	const int iters = 20000;
	for(int rep = 0; rep < 2; ++rep) { k_trip<FENCE, FILL><<<148, warps * 32>>>(o2, iters, cyc); cudaDeviceSynchronize(); }
	long long c0; cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost);
	printf("trip warps=%2d fence=%d fill=%2d FFMA/value: %s, %.1f cycles per trip and SM (all warps), %.1f per warp-trip\n", warps, FENCE, FILL,
	       cudaGetErrorString(cudaGetLastError()), (double) c0 / iters / warps, (double) c0 / iters);
}

int main() {
	uint32_t* d; cudaMalloc(&d, 12 * 4 * 32 * 8 * 4);
	cudaMemset(d, 0xff, 12 * 4 * 32 * 8 * 4);
	k_shapes<<<1, 128>>>(d);
	cudaError_t e = cudaDeviceSynchronize();
	printf("k_shapes: %s\n", cudaGetErrorString(e));
	static uint32_t h[12 * 4 * 32 * 8];
	cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
	const char* names[12] = {"ld 32x32b.x4", "ld 16x64b.x4 lanes0-15", "ld 16x64b.x4 lanes16-31", "ld 16x128b.x2", "ld 16x256b.x1", "ld 16x256b.x2", "ld 16x128b.x4 lanes16-31",
	                         "ld 16x32bx2.x4 imm16", "st 16x256b.x2 -> ld 32x32b cols 32-39", "... cols 40-47", "st 16x128b.x4 -> ld 32x32b cols 48-55", "... cols 56-63"};
	for(int t = 0; t < 12; ++t) {
		printf("== test %d: %s (value = lane*1000+col, or thread*10+reg (+50000 for the upper-16-lane store))\n", t, names[t]);
		for(int w = 0; w < 2; ++w) for(int l = 0; l < 32; ++l) {
			printf("w%d t%02d:", w, l);
			for(int i = 0; i < 8; ++i) { uint32_t v = h[((t * 4 + w) * 32 + l) * 8 + i]; if(v != 0xFFFFFFFFu) printf(" %6u", v); }
			printf("\n");
		}
	}
	long long* cyc; cudaMalloc(&cyc, 148 * 8);
	uint32_t* o2; cudaMalloc(&o2, 148 * 640 * 4);
	const int iters = 20000;
	for(int warps : {4, 8, 20}) {
		for(int mode = 1; mode <= 3; ++mode) {
			cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
			for(int rep = 0; rep < 2; ++rep) {
				cudaEventRecord(a);
				if(mode == 1) k_bw<1><<<148, warps * 32, 32768>>>(o2, iters, cyc);
				if(mode == 2) k_bw<2><<<148, warps * 32, 32768>>>(o2, iters, cyc);
				if(mode == 3) k_bw<3><<<148, warps * 32, 32768>>>(o2, iters, cyc);
				cudaEventRecord(b); e = cudaDeviceSynchronize();
			}
			float ms; cudaEventElapsedTime(&ms, a, b);
			long long c0; cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost);
			// per iteration and warp: TMEM 4 x 512 B, shared 4 x 512 B
			const double bytes = (double) iters * warps * 2048.0;
			printf("bw warps=%2d mode=%d (%s): %s %.3f ms, %lld cycles, %.1f B/clk/SM per memory kind\n", warps, mode, mode == 1 ? "tmem" : mode == 2 ? "smem" : "both",
			       cudaGetErrorString(e), ms, c0, bytes / (double) c0);
		}
	}
	for(int warps : {1, 4, 20}) {
		run_trip<1, 0>(o2, cyc, warps); run_trip<0, 0>(o2, cyc, warps);
		run_trip<1, 8>(o2, cyc, warps); run_trip<0, 8>(o2, cyc, warps);
	}
	return 0;
}
