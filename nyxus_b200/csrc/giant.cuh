// nyxus_b200 -- the large-ROI ("giant") path: one ROI at a time, the whole GPU per ROI.
//
// ROIs whose AABB exceeds the shared-memory tile classes (> 40000 cells: whole-slide single-ROI masks, tissue
// regions, BASELINE configs[2]) cannot be handled by the CTA-per-ROI kernels of kernels.cuh.  Here the ROI's binned
// tile lives in global memory (same cell code as the shared-memory tile: level | CELL_NZ | CELL_IN), every family is
// a grid-wide BUILD kernel (privatised shared-memory histograms per CTA, flushed once into dense / hashed global
// matrices with integer atomics => order independent, deterministic) followed by a small REDUCE kernel that sweeps the
// finished matrix in a fixed order (FP64, fixed trees => deterministic).  With >= 40000 cells per ROI a dense sweep over
// an Ng x Ng / Ng x Nr matrix is cheap next to the build, so the reducers are the textbook dense formulation; the
// per-ROI "owner computes" trick of the small path is not needed.
//
// Semantics are the ones documented in glcm_body.h / glrlm_body.h / zones_body.h / nbr_body.h / intensity_body.h
// (reference file:line cited there); only the execution shape differs.
#pragma once
#include "kernels.cuh"

namespace nyx
{

constexpr int BG_T = 256;		// threads per CTA, every giant kernel (cta_sum supports <= NW_MAX warps)
constexpr int BG_NACC = 32;		// u64 accumulators per family
constexpr int BG_NPART = 16;		// f64 partial sums per CTA
constexpr size_t BG_SMEM_WORDS = 40960;	// shared-memory budget (32-bit words) of the privatised histograms (160 KB)

// contiguous chunk of the tile's cells for this CTA, walked with an incrementally maintained (row, col)
struct CellIter
{
	unsigned long long i, end;
	int row, col;
	__device__ __forceinline__ void init (unsigned long long ncell, int w)
	{
		const unsigned long long per = (ncell + gridDim.x - 1) / gridDim.x;
		const unsigned long long b = per * blockIdx.x;
		end = b + per < ncell ? b + per : ncell;
		i = b + threadIdx.x;
		row = (int)(i / (unsigned long long)w);
		col = (int)(i - (unsigned long long)row * (unsigned long long)w);
	}
	__device__ __forceinline__ bool valid() const { return i < end; }
	__device__ __forceinline__ void next (int w)
	{
		i += blockDim.x;
		col += (int)blockDim.x;
		if (col >= w)
		{
			const int q = col / w;
			row += q;
			col -= q * w;
		}
	}
};

__device__ __forceinline__ Cta big_cta (double* red)
{
	return Cta { (int)threadIdx.x, (int)blockDim.x, red };
}

// per-CTA f64 partial sums: part[blockIdx.x][N] (summed later in block order => deterministic)
template <int N> __device__ __forceinline__ void big_store_partials (const Cta& c, double (&v)[N], double* part)
{
	cta_sum<N> (c, v);
	if (c.tid == 0)
		for (int k = 0; k < N; k++)
			part[(size_t)blockIdx.x * BG_NPART + k] = v[k];
}
template <int N> __device__ __forceinline__ void big_load_partials (const Cta& c, const double* part, int nblk, double (&v)[N], double* sh)
{
	if (c.tid < N)
	{
		double s = 0;
		for (int b = 0; b < nblk; b++)
			s += part[(size_t)b * BG_NPART + c.tid];
		sh[c.tid] = s;
	}
	__syncthreads();
	for (int k = 0; k < N; k++)
		v[k] = sh[k];
	__syncthreads();
}
template <int N> __device__ __forceinline__ void big_add_acc (const Cta& c, unsigned long long (&q)[N], unsigned long long* acc)
{
	cta_sum_u64<N> (c, q);
	if (c.tid == 0)
		for (int k = 0; k < N; k++)
			if (q[k])
				atomicAdd (acc + k, q[k]);
}

__global__ void k_big_fill (double* out, int n, double v)
{
	for (int i = threadIdx.x; i < n; i += blockDim.x)
		out[i] = v;
}

// gather the records of a class list into a contiguous array (one D2H for the host loop)
__global__ void k_gather_rois (const RoiRec* rois, const uint32_t* list, uint32_t n, RoiRec* out, uint32_t* idx_out)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n)
	{
		out[i] = rois[list[i]];
		idx_out[i] = list[i];
	}
}

// ------------------------------------------------------------------------------------------------
// tile materialisation in global memory (image_matrix.h:284-304); cell2: NGLDM's to_grayscale tile (ngldm.cpp:42,126)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BG_T) k_big_tile (ImageView im, RoiRec r, Params prm, uint32_t* __restrict__ cell, uint32_t* __restrict__ cell2)
{
	Binner B;
	B.init (prm.ibsi ? 0 : prm.greydepth, r.imin, r.imax);
	const unsigned long long ncell = (unsigned long long)r.w * (unsigned long long)r.h;
	CellIter it;
	for (it.init (ncell, r.w); it.valid(); it.next (r.w))
	{
		const size_t g = (size_t)(r.y0 + it.row) * im.W + (size_t)(r.x0 + it.col);
		const bool in = im.label[g] == r.label;
		const uint32_t v = in ? load_inten (im, g) : 0u;
		if (cell)
		{
			uint32_t code = B.bin (v) & CELL_LEVEL;
			if (v) code |= CELL_NZ;
			if (in) code |= CELL_IN;
			cell[it.i] = code;
		}
		if (cell2)
			cell2[it.i] = in ? ((to_grayscale (v, 0, r.imax, (uint32_t)prm.greydepth, prm.ibsi != 0) & CELL_LEVEL) | CELL_IN) : 0u;
	}
}

// ------------------------------------------------------------------------------------------------
// GLCM
// ------------------------------------------------------------------------------------------------
// Build: per CTA a BAND of the joint matrix around the diagonal in shared memory (natural images put almost every pair
// there; the full matrix when it fits), pairs outside the band go straight to the global matrix.  BW = odd band width.
__global__ void __launch_bounds__(BG_T) k_big_glcm_build (const uint32_t* __restrict__ cell, int w, int h, int Ng, int off, int sym,
	int BW, uint32_t* joint)
{
	extern __shared__ __align__(16) uint32_t sm[];
	const int B = BW >> 1;
	const unsigned long long ncell = (unsigned long long)w * (unsigned long long)h;
	const size_t nn = (size_t)Ng * Ng;
	for (int a = 0; a < 4; a++)
	{
		const int dx = a == 2 ? 0 : (a == 3 ? -off : off), dy = a == 0 ? 0 : off;	// glcm.cpp:234-251
		uint32_t* J = joint + (size_t)a * nn;
		for (int k = threadIdx.x; k < Ng * BW; k += blockDim.x)
			sm[k] = 0;
		__syncthreads();
		CellIter it;
		for (it.init (ncell, w); it.valid(); it.next (w))
		{
			const uint32_t cc = cell[it.i];
			if (!(cc & CELL_NZ))
				continue;
			const int r2 = it.row + dy, c2 = it.col + dx;
			if (r2 < 0 || r2 >= h || c2 < 0 || c2 >= w)
				continue;
			const uint32_t nc = cell[(unsigned long long)r2 * w + c2];
			if (!(nc & CELL_NZ))
				continue;
			const int b = (int)(cc & CELL_LEVEL) - 1, av = (int)(nc & CELL_LEVEL) - 1;
			const int d = av - b + B;
			if ((unsigned)d < (unsigned)BW)
				atomicAdd (&sm[b * BW + d], 1u);
			else
				atomicAdd (&J[(size_t)b * Ng + av], 1u);
			if (sym)
			{
				const int d2 = b - av + B;
				if ((unsigned)d2 < (unsigned)BW)
					atomicAdd (&sm[av * BW + d2], 1u);
				else
					atomicAdd (&J[(size_t)av * Ng + b], 1u);
			}
		}
		__syncthreads();
		for (int k = threadIdx.x; k < Ng * BW; k += blockDim.x)
		{
			const uint32_t v = sm[k];
			if (v)
			{
				const int b = k / BW, av = b + (k - b * BW) - B;
				atomicAdd (&J[(size_t)b * Ng + av], v);
			}
		}
		__syncthreads();
	}
}

// shared-memory words of the dense GLCM reducer
inline size_t big_glcm_reduce_words (int Ng)
{
	size_t ints = (size_t)Ng * 5;
	ints += ints & 1;
	return ints + (size_t)Ng * 4 + (size_t)Ng + (Ng & 1) + 2;
}

// Reduce: one CTA per angle sweeps the dense joint matrix (row = centre level, column = neighbour level).
__global__ void __launch_bounds__(BG_T) k_big_glcm_reduce (const uint32_t* __restrict__ joint, int Ng, RoiRec r, Params prm, double* out)
{
	extern __shared__ __align__(16) uint32_t sm[];
	__shared__ double red[RED_MAX * NW_MAX];
	__shared__ double sh_m[M_N];
	__shared__ int sh_n[2];
	const Cta c = big_cta (red);
	const int a = blockIdx.x;
	const size_t nn = (size_t)Ng * Ng;
	const uint32_t* J = joint + (size_t)a * nn;
	uint32_t* Rm = sm;
	uint32_t* Cm = Rm + Ng;
	uint32_t* SPM = Cm + Ng;
	uint32_t* SPP = SPM + Ng;
	size_t ints = (size_t)Ng * 5;
	ints += ints & 1;
	double* px = (double*)(sm + ints);
	double* py = px + Ng;
	uint16_t* lx = (uint16_t*)(py + Ng);
	uint16_t* ly = lx + Ng;
	for (int k = c.tid; k < 5 * Ng; k += c.nthr)
		Rm[k] = 0;
	__syncthreads();

	// ---- sweep 1: marginals + integer moments
	enum { Q_N, Q_R, Q_C, Q_D, Q_D2, Q_S, Q_RC, Q_ASM, Q_CNT };
	unsigned long long q[Q_CNT];
	for (int k = 0; k < Q_CNT; k++)
		q[k] = 0;
	unsigned long long jmax = 0;
	for (int b = 0; b < Ng; b++)
		for (int av = c.tid; av < Ng; av += c.nthr)
		{
			const unsigned long long cnt = J[(size_t)b * Ng + av];
			if (!cnt)
				continue;
			const int d = av > b ? av - b : b - av;
			atomicAdd (Rm + b, (uint32_t)cnt);
			atomicAdd (Cm + av, (uint32_t)cnt);
			atomicAdd (SPM + d, (uint32_t)cnt);
			atomicAdd (SPP + av + b, (uint32_t)cnt);
			const unsigned long long lb = (unsigned long long)(b + 1), la = (unsigned long long)(av + 1), ld = (unsigned long long)d;
			q[Q_N] += cnt;
			q[Q_R] += cnt * lb;
			q[Q_C] += cnt * la;
			q[Q_D] += cnt * ld;
			q[Q_D2] += cnt * ld * ld;
			q[Q_S] += cnt * (lb + la);
			q[Q_RC] += cnt * lb * la;
			q[Q_ASM] += cnt * cnt;
			jmax = cnt > jmax ? cnt : jmax;
		}
	__syncthreads();
	cta_sum_u64<Q_CNT> (c, q);
	jmax = cta_max_u64 (c, jmax);
	if (!q[Q_N])	// glcm.cpp:260-295
	{
		for (int f = c.tid; f < G_NANGLED; f += c.nthr)
			out[f * 4 + a] = prm.softnan;
		return;
	}
	const double N = (double)q[Q_N];
	const double mr = (double)q[Q_R] / N, mc = (double)q[Q_C] / N, diffavg = (double)q[Q_D] / N;
	const double NgD = prm.ibsi ? (double)r.imax : (double)Ng;

	// ---- marginal sweep (warp 0, exactly the small path's sweep)
	if (c.tid < 32)
	{
		const Cta g { c.tid, 32, nullptr, 1 };
		double m[M_N];
		for (int k = 0; k < M_N; k++)
			m[k] = 0;
		int nx = 0, ny = 0;
		for (int base = 0; base < 2 * Ng; base += 32)
		{
			const int k = base + g.tid;
			uint32_t cv = 0, rv = 0, dv = 0, sv = 0;
			if (k < Ng)
			{
				cv = Cm[k];
				rv = Rm[k];
				dv = SPM[k];
			}
			if (k < 2 * Ng)
				sv = SPP[k];
			glcm_marginal_terms (k, Ng, cv, rv, dv, sv, N, mr, mc, diffavg, NgD, m, px, py, (double)(k + 1), (double)k, true);
			const unsigned bx = g.ballot (cv != 0), by = g.ballot (rv != 0);
			if (cv)
				lx[nx + popc32 (bx & g.lanemask_lt())] = (uint16_t)k;
			if (rv)
				ly[ny + popc32 (by & g.lanemask_lt())] = (uint16_t)k;
			nx += popc32 (bx);
			ny += popc32 (by);
		}
		cta_sum<M_N> (g, m);
		if (g.tid == 0)
		{
			for (int k = 0; k < M_N; k++)
				sh_m[k] = m[k];
			sh_n[0] = nx;
			sh_n[1] = ny;
		}
	}
	__syncthreads();
	double m[M_N];
	for (int k = 0; k < M_N; k++)
		m[k] = sh_m[k];
	const int nx = sh_n[0], ny = sh_n[1];

	// ---- sweep 2: the joint-cell terms
	double jd[3] = { 0, 0, 0 };	// sum cnt*log(p), sum cnt*log(px*py), sum cnt*(r-mr)(c-mc)
	for (int b = 0; b < Ng; b++)
		for (int av = c.tid; av < Ng; av += c.nthr)
		{
			const unsigned long long cnt = J[(size_t)b * Ng + av];
			if (!cnt)
				continue;
			const double cd = (double)cnt, p = cd / N;
			jd[0] += cd * fast_log10 (dadd (p, GLCM_EPSILON));					// glcm.cpp:868, 734-735, 1160-1161
			jd[1] += cd * fast_log10 (dadd (dmul (px[av], py[b]), GLCM_EPSILON));			// :869
			jd[2] += cd * (((double)(b + 1) - mr) * ((double)(av + 1) - mc));			// :633
		}
	cta_sum<3> (c, jd);
	// ---- HXY2 over the outer product of the non-empty marginals (glcm.cpp:905-911)
	double hxy2[1] = { 0 };
	for (int ii = 0; ii < nx; ii++)
	{
		const double pxi = px[lx[ii]];
		for (int jj = c.tid; jj < ny; jj += c.nthr)
		{
			const double qv = dmul (pxi, py[ly[jj]]);
			hxy2[0] += dmul (qv, fast_log10 (dadd (qv, GLCM_EPSILON)));
		}
	}
	cta_sum<1> (c, hxy2);
	double f[G_NANGLED];
	glcm_finish (f, m, N, mr, diffavg, (double)q[Q_ASM], (double)q[Q_RC], (double)q[Q_D2], (double)q[Q_S], (double)jmax,
		jd[0] / N * INV_LOG10_2, jd[1] / N * INV_LOG10_2, hxy2[0] * INV_LOG10_2, m[M_HX] * INV_LOG10_2, jd[2] / N,
		nx == 1 || ny == 1, true, prm.softnan, diffavg);
	if (c.tid == 0)
		for (int k = 0; k < G_NANGLED; k++)
			out[k * 4 + a] = f[k];
}

// _AVE columns of GLCM (glcm.cpp:1205-1214) and GLRLM (glrlm.cpp:336-351, 902-909)
__global__ void k_big_ave (double* out, int is_glcm)
{
	const int ave_order[G_NAVE] = NYX_GLCM_AVE_ORDER;
	const int n = is_glcm ? G_NAVE : R_NANGLED, base = is_glcm ? G_NANGLED * 4 : R_NANGLED * 4;
	for (int k = threadIdx.x; k < n; k += blockDim.x)
	{
		const double* v = out + (is_glcm ? ave_order[k] : k) * 4;
		out[base + k] = (((v[0] + v[1]) + v[2]) + v[3]) / 4.0;
	}
}

// ------------------------------------------------------------------------------------------------
// GLRLM: dense P[4][Lmax][Nr] in global memory; runs of length <= SL are counted in a per-CTA shared table first
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BG_T) k_big_glrlm_build (const uint32_t* __restrict__ cell, int w, int h, int Lmax, int Nr, int SL,
	uint32_t* P, unsigned long long* acc)
{
	extern __shared__ __align__(16) uint32_t sm[];	// [4][Lmax][SL]
	__shared__ double red[RED_MAX * NW_MAX];
	const Cta c = big_cta (red);
	const unsigned long long ncell = (unsigned long long)w * (unsigned long long)h;
	const size_t nsm = (size_t)4 * Lmax * SL;
	for (size_t k = threadIdx.x; k < nsm; k += blockDim.x)
		sm[k] = 0;
	__syncthreads();
	unsigned long long np[1] = { 0 };
	uint32_t maxlen[4] = { 0, 0, 0, 0 };	// longest run per angle: the reducer sweeps only the populated run lengths
	CellIter it;
	for (it.init (ncell, w); it.valid(); it.next (w))
	{
		const uint32_t cc = cell[it.i];
		if (cc & CELL_NZ)
			np[0]++;						// glrlm.cpp:198-204
		const uint32_t lv = cc & CELL_LEVEL;
		if (!lv)
			continue;
#pragma unroll
		for (int a = 0; a < 4; a++)
		{
			const int dx = a == 2 ? 0 : (a == 3 ? -1 : 1), dy = a == 0 ? 0 : 1;	// E, SE, S, SW (glrlm.cpp:129-175)
			const int pr = it.row - dy, pc = it.col - dx;
			if (pr >= 0 && pc >= 0 && pc < w && (cell[(unsigned long long)pr * w + pc] & CELL_LEVEL) == lv)
				continue;					// not a run start
			int len = 1, y = it.row + dy, x = it.col + dx;
			while (y < h && x >= 0 && x < w && (cell[(unsigned long long)y * w + x] & CELL_LEVEL) == lv)
			{
				len++;
				y += dy;
				x += dx;
			}
			if (len <= SL)
				atomicAdd (&sm[((size_t)a * Lmax + (lv - 1)) * SL + (len - 1)], 1u);
			else
				atomicAdd (&P[((size_t)a * Lmax + (lv - 1)) * Nr + (len - 1)], 1u);
			maxlen[a] = (uint32_t)len > maxlen[a] ? (uint32_t)len : maxlen[a];
		}
	}
#pragma unroll
	for (int a = 0; a < 4; a++)
	{
		const uint32_t m = __reduce_max_sync (0xffffffffu, maxlen[a]);
		if ((threadIdx.x & 31) == 0 && m)
			atomicMax (&acc[BG_NACC + a], (unsigned long long)m);
	}
	__syncthreads();
	for (size_t k = threadIdx.x; k < nsm; k += blockDim.x)
	{
		const uint32_t v = sm[k];
		if (v)
		{
			const size_t al = k / SL;
			atomicAdd (&P[al * Nr + (k - al * SL)], v);
		}
	}
	big_add_acc<1> (c, np, acc);
}

__global__ void __launch_bounds__(BG_T) k_big_glrlm_reduce (const uint32_t* __restrict__ Pall, int Lmax, int Nr, const unsigned long long* acc, double* out)
{
	extern __shared__ __align__(16) uint32_t sm[];	// Ri[Lmax] + Cj[Nr]
	__shared__ double red[RED_MAX * NW_MAX];
	const Cta c = big_cta (red);
	const int a = blockIdx.x;
	const uint32_t* P = Pall + (size_t)a * Lmax * Nr;
	uint32_t* Ri = sm;
	uint32_t* Cj = Ri + Lmax;
	for (int k = c.tid; k < Lmax + Nr; k += c.nthr)
		sm[k] = 0;
	__syncthreads();
	// columns beyond the longest run of this angle are empty (k_big_glrlm_build records it): a 16k x 16k slide has
	// Nr = 16384 columns and runs of a few dozen pixels
	const int Nre = (int)min ((unsigned long long)Nr, acc[BG_NACC + a]);
	enum { Q_SUM, Q_LRE, Q_MUI, Q_MUJ, Q_HGL, Q_N };
	unsigned long long q[Q_N];
	for (int k = 0; k < Q_N; k++)
		q[k] = 0;
	enum { A_SRE, A_LGL, A_SRL, A_SRH, A_LRL, A_LRH, A_N };
	double s[A_N];
	for (int k = 0; k < A_N; k++)
		s[k] = 0;
	for (int l = 0; l < Lmax; l++)
		for (int j = c.tid; j < Nre; j += c.nthr)
		{
			const unsigned long long cnt = P[(size_t)l * Nr + j];
			if (!cnt)
				continue;
			const uint32_t lv = (uint32_t)(l + 1);
			const int len = j + 1;
			atomicAdd (Ri + l, (uint32_t)cnt);
			atomicAdd (Cj + j, (uint32_t)cnt);
			const uint32_t i2 = lv * lv, j2 = (uint32_t)(len * len), i2j2 = lv * lv * (uint32_t)len * (uint32_t)len;	// 32-bit wrap
			const double cd = (double)cnt;
			q[Q_SUM] += cnt;
			q[Q_LRE] += cnt * (unsigned long long)len * (unsigned long long)len;	// glrlm.cpp:409-415
			q[Q_MUI] += cnt * lv;							// :598-605
			q[Q_MUJ] += cnt * (unsigned long long)len;				// :640-646
			q[Q_HGL] += cnt * (unsigned long long)i2;				// :744-751
			s[A_SRE] += cd / (double)(len * len);					// :381-383
			s[A_LGL] += cd / (double)i2;						// :712-719
			s[A_SRL] += cd / (double)i2j2;						// :781
			s[A_SRH] += cd * (double)i2 / (double)(int)j2;				// :813
			s[A_LRL] += cd * (double)(int)j2 / (double)i2;				// :844
			s[A_LRH] += cd * (double)i2j2;						// :876
		}
	__syncthreads();
	cta_sum_u64<Q_N> (c, q);
	cta_sum<A_N> (c, s);
	if (q[Q_SUM] == 0)	// glrlm.cpp:363-366 etc.
	{
		for (int f = c.tid; f < R_NANGLED; f += c.nthr)
			out[f * 4 + a] = 0.0;
		return;
	}
	const double sum = (double)q[Q_SUM], Np = (double)acc[0];
	const double mui = (double)q[Q_MUI] / sum, muj = (double)q[Q_MUJ] / sum;
	unsigned long long q2[2] = { 0, 0 };
	for (int k = c.tid; k < Lmax; k += c.nthr)
		q2[0] += (unsigned long long)Ri[k] * Ri[k];			// :441-448
	for (int k = c.tid; k < Nre; k += c.nthr)
		q2[1] += (unsigned long long)Cj[k] * Cj[k];			// :509-517
	double t[3] = { 0, 0, 0 };
	for (int l = 0; l < Lmax; l++)
		for (int j = c.tid; j < Nre; j += c.nthr)
		{
			const unsigned long long cnt = P[(size_t)l * Nr + j];
			if (!cnt)
				continue;
			const double cd = (double)cnt, di = (double)(l + 1) - mui, dj = (double)(j + 1) - muj;
			t[0] += cd * di * di;								// :608-616
			t[1] += cd * dj * dj;								// :649-656
			t[2] += cd * fast_log10 (dadd (cd / sum, EPS_ZONES));				// :684-685
		}
	cta_sum_u64<2> (c, q2);
	cta_sum<3> (c, t);
	if (c.tid == 0)
	{
		double f[R_NANGLED];
		f[R_SRE] = s[A_SRE] / sum;
		f[R_LRE] = (double)q[Q_LRE] / sum;
		f[R_GLN] = (double)q2[0] / sum;
		f[R_GLNN] = (double)q2[0] / (sum * sum);
		f[R_RLN] = (double)q2[1] / sum;
		f[R_RLNN] = (double)q2[1] / (sum * sum);
		f[R_RP] = sum / Np;					// :573
		f[R_GLV] = t[0] / sum;
		f[R_RV] = t[1] / sum;
		f[R_RE] = -(t[2] / sum) * (1.0 / LOG10_2);
		f[R_LGLRE] = s[A_LGL] / sum;
		f[R_HGLRE] = (double)q[Q_HGL] / sum;
		f[R_SRLGLE] = s[A_SRL] / sum;
		f[R_SRHGLE] = s[A_SRH] / sum;
		f[R_LRLGLE] = s[A_LRL] / sum;
		f[R_LRHGLE] = s[A_LRH] / sum;
		for (int k = 0; k < R_NANGLED; k++)
			out[k * 4 + a] = f[k];
	}
}


// ------------------------------------------------------------------------------------------------
// open-addressing table with 64-bit keys (zone sizes of a giant ROI do not fit the 16-bit minor key of zones_body.h)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t htab64_mix (unsigned long long k)
{
	k ^= k >> 33;
	k *= 0xff51afd7ed558ccdull;
	k ^= k >> 33;
	return (uint32_t)k;
}
__device__ __forceinline__ void htab64_add (unsigned long long* keys, uint32_t* cnts, uint32_t mask, unsigned long long key, uint32_t v)
{
	uint32_t slot = htab64_mix (key) & mask;
	for (;;)
	{
		unsigned long long k = ((volatile unsigned long long*)keys)[slot];
		if (k == 0)
			k = atomicCAS (keys + slot, 0ull, key);
		if (k == 0 || k == key)
			break;
		slot = (slot + 1) & mask;
	}
	atomicAdd (cnts + slot, v);
}

// ------------------------------------------------------------------------------------------------
// GLSZM: forward-reachability labels  label(p) = min( idx(p), label(q) : q in {W,NW,N,NE}, level(q) == level(p) )
// (zones_body.h).  Rows depend on the row above, so a giant tile is cut into bands of R rows; every band is solved
// top-down by one CTA taking the last row of the band above as it currently is, and bands whose input row changed are
// solved again (chaotic relaxation: labels only ever decrease towards the unique fixed point of the recurrence, every
// stored value is the index of a true ancestor).  Natural images converge in 2-3 sweeps; a zone spanning k bands needs
// k sweeps of the bands it crosses only.
// ------------------------------------------------------------------------------------------------
__global__ void k_big_iota (uint32_t* a, unsigned long long n)
{
	for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x)
		a[i] = (uint32_t)i;
}

constexpr int GZ_E = 4;			// consecutive cells per thread
constexpr int GZ_CH = BG_T * GZ_E;	// cells per chunk of a row

__global__ void __launch_bounds__(BG_T) k_big_glszm_pass (const uint32_t* __restrict__ cell, int w, int h, int R, uint32_t* lab,
	const uint32_t* dirty_in, uint32_t* dirty_out, uint32_t* any_change)
{
	const int band = blockIdx.x;
	if (!dirty_in[band])
		return;
	__shared__ uint32_t s_tail[BG_T / 32], s_head[BG_T / 32], s_carry[BG_T / 32], s_chunk_carry, s_changed;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int y0 = band * R, y1 = min (h, y0 + R);
	bool changed_any = false, changed_last = false;
	for (int y = y0; y < y1; y++)
	{
		bool changed_row = false;
		if (threadIdx.x == 0)
			s_chunk_carry = 0xFFFFFFFFu;
		__syncthreads();
		for (int xc = 0; xc < w; xc += GZ_CH)
		{
			const int xb = xc + threadIdx.x * GZ_E;
			uint32_t lv[GZ_E], m[GZ_E], old[GZ_E];
			bool hd[GZ_E];
			uint32_t left = 0xFFFFFFFFu;	// level of the cell left of this thread's first cell
			if (xb > 0 && xb < w)
				left = cell[(unsigned long long)y * w + xb - 1] & CELL_LEVEL;
#pragma unroll
			for (int k = 0; k < GZ_E; k++)
			{
				const int x = xb + k;
				lv[k] = 0;
				m[k] = 0xFFFFFFFFu;
				old[k] = 0xFFFFFFFFu;
				if (x < w)
				{
					const unsigned long long i = (unsigned long long)y * w + x;
					lv[k] = cell[i] & CELL_LEVEL;
					m[k] = (uint32_t)i;
					if (lv[k])
					{
						old[k] = lab[i];
						if (y > 0)
						{
							const unsigned long long up = i - w;
							if (x > 0 && (cell[up - 1] & CELL_LEVEL) == lv[k]) m[k] = min (m[k], ((volatile uint32_t*)lab)[up - 1]);
							if ((cell[up] & CELL_LEVEL) == lv[k]) m[k] = min (m[k], ((volatile uint32_t*)lab)[up]);
							if (x + 1 < w && (cell[up + 1] & CELL_LEVEL) == lv[k]) m[k] = min (m[k], ((volatile uint32_t*)lab)[up + 1]);
						}
					}
				}
				hd[k] = !lv[k] || lv[k] != (k ? lv[k - 1] : left);
			}
			// running minimum inside the thread
#pragma unroll
			for (int k = 1; k < GZ_E; k++)
				if (!hd[k])
					m[k] = min (m[k], m[k - 1]);
			const bool anyhead = hd[0] | hd[1] | hd[2] | hd[3];
			// segmented inclusive min-scan of the thread tails over the warp
			uint32_t sv = m[GZ_E - 1];
			const unsigned heads = __ballot_sync (0xffffffffu, anyhead);
			const int start = (heads & (0xFFFFFFFFu >> (31 - lane))) ? 31 - __clz (heads & (0xFFFFFFFFu >> (31 - lane))) : -1;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1)
			{
				const uint32_t t = __shfl_up_sync (0xffffffffu, sv, o);
				if (lane - o >= start && lane - o >= 0 && !(start == lane))
					sv = min (sv, t);
			}
			// (a thread that contains a head starts a new segment: its tail must not take values from the left)
			if (lane == 31)
			{
				s_tail[warp] = sv;
				s_head[warp] = heads != 0;
			}
			__syncthreads();
			if (threadIdx.x == 0)
			{
				uint32_t run = s_chunk_carry;
				for (int wq = 0; wq < BG_T / 32; wq++)
				{
					s_carry[wq] = run;	// what flows into warp wq
					run = s_head[wq] ? s_tail[wq] : min (run, s_tail[wq]);
				}
				s_chunk_carry = run;
			}
			__syncthreads();
			// carry into this thread = scan value of the previous lane, or the warp's carry-in when no head lies in between
			uint32_t cin = __shfl_up_sync (0xffffffffu, sv, 1);
			const unsigned before = heads & ((1u << lane) - 1u);	// heads in lanes < lane
			if (lane == 0)
				cin = s_carry[warp];
			else if (!before)
				cin = min (cin, s_carry[warp]);
			// cells before the first head of the thread take the carry
#pragma unroll
			for (int k = 0; k < GZ_E; k++)
			{
				if (hd[k])
					break;
				m[k] = min (m[k], cin);
			}
#pragma unroll
			for (int k = 0; k < GZ_E; k++)
				if (xb + k < w && lv[k] && m[k] != old[k])
				{
					lab[(unsigned long long)y * w + xb + k] = m[k];
					changed_row = true;
				}
			__syncthreads();
		}
		changed_any |= changed_row;
		if (y == y1 - 1)
			changed_last = changed_row;
		__threadfence_block();
		__syncthreads();	// the next row reads this row's labels
	}
	if (threadIdx.x == 0)
		s_changed = 0;
	__syncthreads();
	if (changed_any)
		atomicOr (&s_changed, 1u);
	if (changed_last)
		atomicOr (&s_changed, 2u);
	__syncthreads();
	if (threadIdx.x == 0 && s_changed)
	{
		*any_change = 1;
		dirty_out[band] = 1;			// solve this band once more next sweep (cheap, and robust against the racing reads)
		if ((s_changed & 2u) && y1 < h)
			dirty_out[band + 1] = 1;
	}
}

// zone sizes: one atomic per horizontal stretch of equal labels inside a warp
__global__ void __launch_bounds__(BG_T) k_big_zone_sizes (const uint32_t* __restrict__ cell, const uint32_t* __restrict__ lab, uint32_t* zs,
	unsigned long long ncell, int skipzero_only_ibsi, int ibsi, unsigned long long* acc_np)
{
	__shared__ double red[RED_MAX * NW_MAX];
	const Cta c = big_cta (red);
	const int lane = threadIdx.x & 31;
	const unsigned long long per = (ncell + gridDim.x - 1) / gridDim.x;
	const unsigned long long b0 = per * blockIdx.x, b1 = min (b0 + per, ncell);
	unsigned long long np[1] = { 0 };
	for (unsigned long long base = b0; base < b1; base += blockDim.x)
	{
		const unsigned long long i = base + threadIdx.x;
		uint32_t l = 0xFFFFFFFFu;
		bool on = false;
		if (i < b1)
		{
			const uint32_t lv = cell[i] & CELL_LEVEL;
			on = skipzero_only_ibsi ? !(ibsi && !lv) : lv != 0;
			if (on)
			{
				l = lab[i];
				np[0]++;
			}
		}
		const uint32_t pl = __shfl_up_sync (0xffffffffu, l, 1);
		const bool head = on && (lane == 0 || pl != l);
		const unsigned hm = __ballot_sync (0xffffffffu, head), om = __ballot_sync (0xffffffffu, on);
		if (head)
		{
			// stretch = this lane up to (not including) the next head or the next inactive lane
			const unsigned stop = (hm | ~om) & ~((2u << lane) - 1u);
			const int endl = stop ? __ffs (stop) - 1 : 32;
			atomicAdd (&zs[l], (uint32_t)(endl - lane));
		}
	}
	big_add_acc<1> (c, np, acc_np);
}

constexpr int GZ_SS = 1024;		// zone sizes <= GZ_SS: dense size marginal
constexpr int GZ_SJ = 8;		// zone sizes <= GZ_SJ: dense joint (level, size) table
enum { ZQ_NZ, ZQ_MUJ, ZQ_MUI, ZQ_HGL, ZQ_N };
enum { ZA_SAE, ZA_LAE, ZA_LGL, ZA_SAL, ZA_SAH, ZA_LAL, ZA_LAH, ZA_N };

// words of the dense global tables of GLSZM: Gl[L1] + Gs[SS+1] + Gj[L1*SJ]
__host__ __device__ inline size_t big_glszm_dense_words (int L1) { return (size_t)L1 + GZ_SS + 1 + (size_t)L1 * GZ_SJ; }

// zone pass: seeds (lab == own index) -> marginals, joint table, per-zone sums.  use_sm: the three tables are privatised
// in shared memory (L1 * (SJ+1) + SS + 1 words) and flushed once; otherwise global atomics.
__global__ void __launch_bounds__(BG_T) k_big_glszm_zones (const uint32_t* __restrict__ cell, const uint32_t* __restrict__ lab,
	const uint32_t* __restrict__ zs, unsigned long long ncell, int L1, int use_sm, uint32_t* dense, unsigned long long* hkeys,
	uint32_t* hcnts, uint32_t hmask, unsigned long long* acc, double* part, uint32_t* dbg_keys, uint32_t* dbg_n)
{
	extern __shared__ __align__(16) uint32_t sm[];
	__shared__ double red[RED_MAX * NW_MAX];
	const Cta c = big_cta (red);
	const size_t nd = big_glszm_dense_words (L1);
	uint32_t* T = use_sm ? sm : dense;
	uint32_t* Tl = T;
	uint32_t* Ts = Tl + L1;
	uint32_t* Tj = Ts + GZ_SS + 1;
	if (use_sm)
	{
		for (size_t k = threadIdx.x; k < nd; k += blockDim.x)
			sm[k] = 0;
		__syncthreads();
	}
	unsigned long long q[ZQ_N];
	for (int k = 0; k < ZQ_N; k++)
		q[k] = 0;
	double s[ZA_N];
	for (int k = 0; k < ZA_N; k++)
		s[k] = 0;
	const unsigned long long per = (ncell + gridDim.x - 1) / gridDim.x;
	const unsigned long long b0 = per * blockIdx.x, b1 = min (b0 + per, ncell);
	for (unsigned long long i = b0 + threadIdx.x; i < b1; i += blockDim.x)
	{
		const uint32_t lv = cell[i] & CELL_LEVEL;
		if (!lv || lab[i] != (uint32_t)i)
			continue;
		const uint32_t sz = zs[i];
		atomicAdd (Tl + lv, 1u);
		if (sz <= GZ_SS)
			atomicAdd (Ts + sz, 1u);
		else
			htab64_add (hkeys, hcnts, hmask, (unsigned long long)sz, 1u);
		if (sz <= GZ_SJ)
			atomicAdd (Tj + (size_t)lv * GZ_SJ + (sz - 1), 1u);
		else
			htab64_add (hkeys, hcnts, hmask, ((unsigned long long)lv << 32) | sz, 1u);
		if (dbg_keys)
		{
			const uint32_t k = atomicAdd (dbg_n, 1u);
			dbg_keys[2 * (size_t)k] = lv;
			dbg_keys[2 * (size_t)k + 1] = sz;
		}
		q[ZQ_NZ] += 1;
		q[ZQ_MUJ] += sz;						// glszm.cpp:380
		q[ZQ_MUI] += lv;						// :381
		q[ZQ_HGL] += (unsigned long long)lv * lv;			// :574-578
		const double i2 = (double)lv * (double)lv, j2 = (double)sz * (double)sz;
		const double jj = (double)(int)(sz * sz);			// :424-427, :437-440  `j * j` in int (wraps for j >= 46341)
		s[ZA_SAE] += 1.0 / jj;
		s[ZA_LAE] += jj;
		s[ZA_LGL] += 1.0 / i2;						// :559-563
		s[ZA_LAH] += i2 * j2;						// :372-375
		s[ZA_LAL] += j2 / i2;
		s[ZA_SAH] += i2 / j2;
		s[ZA_SAL] += 1.0 / (i2 * j2);
	}
	__syncthreads();
	if (use_sm)
		for (size_t k = threadIdx.x; k < nd; k += blockDim.x)
		{
			const uint32_t v = sm[k];
			if (v)
				atomicAdd (dense + k, v);
		}
	big_add_acc<ZQ_N> (c, q, acc);
	big_store_partials<ZA_N> (c, s, part);
}

__global__ void __launch_bounds__(BG_T) k_big_glszm_reduce (const uint32_t* __restrict__ dense, int L1, const unsigned long long* __restrict__ hkeys,
	const uint32_t* __restrict__ hcnts, uint32_t hslots, const unsigned long long* acc, const unsigned long long* acc_np,
	const double* part, int nblk, unsigned long long Ns, Params prm, double* out)
{
	__shared__ double red[RED_MAX * NW_MAX];
	__shared__ double sh[BG_NPART];
	const Cta c = big_cta (red);
	const uint32_t* Tl = dense;
	const uint32_t* Ts = Tl + L1;
	const uint32_t* Tj = Ts + GZ_SS + 1;
	const unsigned long long nz = acc[ZQ_NZ];
	if (nz == 0)	// glszm.cpp:244-248
	{
		for (int i = c.tid; i < NCOL_GLSZM; i += c.nthr)
			out[i] = prm.softnan;
		return;
	}
	double s[ZA_N];
	big_load_partials<ZA_N> (c, part, nblk, s, sh);
	const double sum = (double)nz;
	const double mui = (double)acc[ZQ_MUI] / sum, muj = (double)acc[ZQ_MUJ] / sum;
	unsigned long long q2[2] = { 0, 0 };
	double t[3] = { 0, 0, 0 };	// GLV, ZV, ZE
	for (int lv = c.tid; lv < L1; lv += c.nthr)
	{
		const unsigned long long n = Tl[lv];
		if (!n)
			continue;
		q2[0] += n * n;						// :450-454
		const double di = (double)lv - mui;
		t[0] += (double)n * di * di;				// :518-526
	}
	for (int sz = c.tid; sz <= GZ_SS; sz += c.nthr)
	{
		const unsigned long long n = Ts[sz];
		if (!n)
			continue;
		q2[1] += n * n;						// :481-485
		const double dj = (double)sz - muj;
		t[1] += (double)n * dj * dj;				// :535-542
	}
	for (size_t k = c.tid; k < (size_t)L1 * GZ_SJ; k += c.nthr)
	{
		const uint32_t n = Tj[k];
		if (n)
			t[2] += (double)n * fast_log10 (dadd ((double)n / sum, EPS_ZONES));	// :377-378
	}
	for (uint32_t k = c.tid; k < hslots; k += c.nthr)
	{
		const unsigned long long key = hkeys[k];
		if (!key)
			continue;
		const unsigned long long n = hcnts[k];
		if (key >> 32)
			t[2] += (double)n * fast_log10 (dadd ((double)n / sum, EPS_ZONES));
		else
		{
			q2[1] += n * n;
			const double dj = (double)(uint32_t)key - muj;
			t[1] += (double)n * dj * dj;
		}
	}
	cta_sum_u64<2> (c, q2);
	cta_sum<3> (c, t);
	if (c.tid == 0)
	{
		// glszm.cpp:420-429: the SAE loop divides sj[j] by `j * j` evaluated in int for EVERY j <= Ns = w*h; at j = 65536 the
		// product wraps to 0 and 0.0 / 0 = NaN poisons the sum (-> noval at output) whenever the tile has >= 65536 cells
		out[S_SAE] = Ns >= 65536ull ? nan ("") : s[ZA_SAE] / sum;
		out[S_LAE] = s[ZA_LAE] / sum;
		out[S_GLN] = (double)q2[0] / sum;
		out[S_GLNN] = (double)q2[0] / (sum * sum);
		out[S_SZN] = (double)q2[1] / sum;
		out[S_SZNN] = (double)q2[1] / (sum * sum);
		out[S_ZP] = sum / (double)acc_np[0];			// :507-511
		out[S_GLV] = t[0] / sum;
		out[S_ZV] = t[1] / sum;
		out[S_ZE] = -(t[2] / sum) * (1.0 / LOG10_2);
		out[S_LGLZE] = s[ZA_LGL] / sum;
		out[S_HGLZE] = (double)acc[ZQ_HGL] / sum;
		out[S_SALGLE] = s[ZA_SAL] / sum;
		out[S_SAHGLE] = s[ZA_SAH] / sum;
		out[S_LALGLE] = s[ZA_LAL] / sum;
		out[S_LAHGLE] = s[ZA_LAH] / sum;
	}
}


// ------------------------------------------------------------------------------------------------
// GLDZM: 4-connected components of equal level by union-find in global memory (zones_body.h), zone metric = minimum
// distance to the border over the members (gldzm.cpp:306-351)
// ------------------------------------------------------------------------------------------------
// every cell starts hooked to the head of its horizontal same-level run: one warp per row, ballot + carry
__global__ void __launch_bounds__(BG_T) k_big_runheads (const uint32_t* __restrict__ cell, int w, int h, int ibsi, uint32_t* parent, uint32_t* zmin)
{
	const int lane = threadIdx.x & 31;
	const int wpb = blockDim.x >> 5;
	for (int y = blockIdx.x * wpb + (threadIdx.x >> 5); y < h; y += gridDim.x * wpb)
	{
		uint32_t carry_lv = 0xFFFFFFFFu, carry_head = 0;
		for (int x0 = 0; x0 < w; x0 += 32)
		{
			const int x = x0 + lane;
			const unsigned long long i = (unsigned long long)y * w + x;
			uint32_t lv = 0xFFFFFFFEu;
			if (x < w)
				lv = cell[i] & CELL_LEVEL;
			uint32_t left = __shfl_up_sync (0xffffffffu, lv, 1);
			if (lane == 0)
				left = carry_lv;
			const bool excluded = ibsi && !lv;			// zero level: not part of any zone (gldzm.cpp:104-106)
			const bool head = excluded || left != lv;
			const unsigned heads = __ballot_sync (0xffffffffu, head);
			const unsigned le = heads & (0xFFFFFFFFu >> (31 - lane));
			const uint32_t hd = le ? (uint32_t)((unsigned long long)y * w + x0 + (31 - __clz (le))) : carry_head;
			if (x < w)
			{
				parent[i] = hd;
				zmin[i] = 0x7FFFFFFFu;
			}
			carry_lv = __shfl_sync (0xffffffffu, lv, 31);
			carry_head = __shfl_sync (0xffffffffu, hd, 31);
		}
	}
}

__global__ void __launch_bounds__(BG_T) k_big_unite_rows (const uint32_t* __restrict__ cell, int w, int h, int ibsi, uint32_t* parent)
{
	const unsigned long long ncell = (unsigned long long)w * (unsigned long long)h;
	const Cta c { 0, 1, nullptr };
	CellIter it;
	for (it.init (ncell, w); it.valid(); it.next (w))
	{
		if (it.row == 0)
			continue;
		const uint32_t lv = cell[it.i] & CELL_LEVEL;
		if ((ibsi && !lv) || (cell[it.i - w] & CELL_LEVEL) != lv)
			continue;
		if (it.col > 0 && (cell[it.i - 1] & CELL_LEVEL) == lv && (cell[it.i - w - 1] & CELL_LEVEL) == lv)
			continue;	// the left neighbours already joined the same two runs
		uf_unite (c, parent, (uint32_t)it.i, (uint32_t)(it.i - w));
	}
}

// parent[i] <- root(i) (a root is an ancestor of everything below it, so concurrent readers stay correct)
__global__ void __launch_bounds__(BG_T) k_big_flatten (uint32_t* parent, unsigned long long ncell)
{
	for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < ncell; i += (unsigned long long)gridDim.x * blockDim.x)
		parent[i] = uf_find (parent, (uint32_t)i);
}

// IBSI mode only (the MATLAB-binned tile has no zero cells): per row the distance to the nearest zero cell / tile edge
// on the left and on the right, in the reference's convention (zones_body.h gldzm_dist2border):
//   dl = x - max(0, last zero left of x) + 1,  dr = min(w-1, first zero right of x) - x + 1
__global__ void __launch_bounds__(BG_T) k_big_dist_rows (const uint32_t* __restrict__ cell, int w, int h, uint32_t* dmap)
{
	const int lane = threadIdx.x & 31;
	const int wpb = blockDim.x >> 5;
	for (int y = blockIdx.x * wpb + (threadIdx.x >> 5); y < h; y += gridDim.x * wpb)
	{
		const unsigned long long rb = (unsigned long long)y * w;
		int lastz = 0;		// max(0, position of the last zero strictly left of the chunk)
		for (int x0 = 0; x0 < w; x0 += 32)
		{
			const int x = x0 + lane;
			const bool z = x < w && (cell[rb + x] & CELL_LEVEL) == 0;
			const unsigned zm = __ballot_sync (0xffffffffu, z);
			const unsigned lt = zm & ((1u << lane) - 1u);
			const int lz = lt ? x0 + 31 - __clz (lt) : lastz;
			if (x < w)
				dmap[rb + x] = (uint32_t)(x - lz + 1);
			if (zm)
				lastz = x0 + 31 - __clz (zm);
		}
		int nextz = w - 1;
		for (int x0 = ((w - 1) / 32) * 32; x0 >= 0; x0 -= 32)
		{
			const int x = x0 + lane;
			const bool z = x < w && (cell[rb + x] & CELL_LEVEL) == 0;
			const unsigned zm = __ballot_sync (0xffffffffu, z);
			const unsigned gt = lane == 31 ? 0u : (zm & ~((2u << lane) - 1u));
			const int rz = gt ? x0 + __ffs (gt) - 1 : nextz;
			if (x < w)
				dmap[rb + x] = min (dmap[rb + x], (uint32_t)(rz - x + 1));
			if (zm)
				nextz = x0 + __ffs (zm) - 1;
		}
	}
}
// columns: one thread per column, walking down and then up
__global__ void __launch_bounds__(BG_T) k_big_dist_cols (const uint32_t* __restrict__ cell, int w, int h, uint32_t* dmap)
{
	for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < w; x += gridDim.x * blockDim.x)
	{
		int lastz = 0;
		for (int y = 0; y < h; y++)
		{
			const unsigned long long i = (unsigned long long)y * w + x;
			dmap[i] = min (dmap[i], (uint32_t)(y - lastz + 1));
			if ((cell[i] & CELL_LEVEL) == 0)
				lastz = y;
		}
		int nextz = h - 1;
		for (int y = h - 1; y >= 0; y--)
		{
			const unsigned long long i = (unsigned long long)y * w + x;
			dmap[i] = min (dmap[i], (uint32_t)(nextz - y + 1));
			if ((cell[i] & CELL_LEVEL) == 0)
				nextz = y;
		}
	}
}

// zone metric: atomicMin of the member distances at the root (root[] flattened)
__global__ void __launch_bounds__(BG_T) k_big_zone_metric (const uint32_t* __restrict__ cell, const uint32_t* __restrict__ root, const uint32_t* __restrict__ dmap,
	int w, int h, int ibsi, uint32_t* zmin)
{
	const unsigned long long ncell = (unsigned long long)w * (unsigned long long)h;
	CellIter it;
	for (it.init (ncell, w); it.valid(); it.next (w))
	{
		const uint32_t lv = cell[it.i] & CELL_LEVEL;
		if (ibsi && !lv)
			continue;
		uint32_t d;
		if (dmap)
			d = dmap[it.i];
		else
		{
			const int dl = it.col + 1, dr = w - it.col, dt = it.row + 1, db = h - it.row;
			d = (uint32_t)min (min (dl, dr), min (dt, db));
		}
		const uint32_t rt = root[it.i];
		if (d < ((volatile uint32_t*)zmin)[rt])
			atomicMin (zmin + rt, d);
	}
}

constexpr int GD_SD = 8;	// zone metrics <= GD_SD: joint counts privatised in shared memory
enum { DQ_NS, DQ_LDE, DQ_ZDM, DQ_GLM, DQ_HGL, DQ_N };
enum { DA_SDE, DA_LGL, DA_SDL, DA_SDH, DA_LDL, DA_LDH, DA_N };

// zones (= roots) -> dense P[L1][Nd] (+ per-zone sums).  sm: Hj[L1][SD] joint counts of the small metrics.
__global__ void __launch_bounds__(BG_T) k_big_gldzm_zones (const uint32_t* __restrict__ cell, const uint32_t* __restrict__ root, const uint32_t* __restrict__ zmin,
	unsigned long long ncell, int ibsi, int L1, int Nd, int use_sm, uint32_t* P, unsigned long long* acc, double* part,
	uint32_t* dbg_keys, uint32_t* dbg_n)
{
	extern __shared__ __align__(16) uint32_t sm[];
	__shared__ double red[RED_MAX * NW_MAX];
	const Cta c = big_cta (red);
	const size_t nsm = use_sm ? (size_t)L1 * GD_SD : 0;
	for (size_t k = threadIdx.x; k < nsm; k += blockDim.x)
		sm[k] = 0;
	__syncthreads();
	unsigned long long q[DQ_N];
	for (int k = 0; k < DQ_N; k++)
		q[k] = 0;
	double s[DA_N];
	for (int k = 0; k < DA_N; k++)
		s[k] = 0;
	const unsigned long long per = (ncell + gridDim.x - 1) / gridDim.x;
	const unsigned long long b0 = per * blockIdx.x, b1 = min (b0 + per, ncell);
	for (unsigned long long i = b0 + threadIdx.x; i < b1; i += blockDim.x)
	{
		const uint32_t lv = cell[i] & CELL_LEVEL;
		if ((ibsi && !lv) || root[i] != (uint32_t)i)
			continue;
		const uint32_t dd = zmin[i];
		if (use_sm && dd <= GD_SD)
			atomicAdd (&sm[(size_t)lv * GD_SD + (dd - 1)], 1u);
		else
			atomicAdd (&P[(size_t)lv * Nd + dd], 1u);
		if (dbg_keys)
		{
			const uint32_t k = atomicAdd (dbg_n, 1u);
			dbg_keys[2 * (size_t)k] = lv;
			dbg_keys[2 * (size_t)k + 1] = dd;
		}
		q[DQ_NS] += 1;
		q[DQ_LDE] += (unsigned long long)dd * dd;			// gldzm.cpp:403
		q[DQ_ZDM] += dd;						// :445
		q[DQ_GLM] += lv;						// :444
		q[DQ_HGL] += (unsigned long long)lv * lv;			// :422
		const double g = (double)lv, d = (double)dd;
		s[DA_SDE] += 1.0 / d / d;					// :402
		s[DA_LGL] += 1.0 / (g * g);					// :421
		s[DA_SDL] += 1.0 / g / g / d / d;				// :440
		s[DA_SDH] += g * g / d / d;					// :441
		s[DA_LDL] += d * d / g / g;					// :442
		s[DA_LDH] += g * g * d * d;					// :443
	}
	__syncthreads();
	for (size_t k = threadIdx.x; k < nsm; k += blockDim.x)
	{
		const uint32_t v = sm[k];
		if (v)
		{
			const size_t lv = k / GD_SD;
			atomicAdd (&P[lv * Nd + (k - lv * GD_SD) + 1], v);
		}
	}
	big_add_acc<DQ_N> (c, q, acc);
	big_store_partials<DA_N> (c, s, part);
}

__global__ void __launch_bounds__(BG_T) k_big_gldzm_reduce (const uint32_t* __restrict__ P, int L1, int Nd, const unsigned long long* acc,
	const double* part, int nblk, RoiRec r, Params prm, double* out)
{
	extern __shared__ __align__(16) uint32_t sm[];	// Cd[Nd]
	__shared__ double red[RED_MAX * NW_MAX];
	__shared__ double sh[BG_NPART];
	const Cta c = big_cta (red);
	uint32_t* Cd = sm;
	for (int k = c.tid; k < Nd; k += c.nthr)
		Cd[k] = 0;
	__syncthreads();
	double s[DA_N];
	big_load_partials<DA_N> (c, part, nblk, s, sh);
	const double Ns = (double)acc[DQ_NS], Nv = (double)r.area;
	const double glm = (double)acc[DQ_GLM] / Ns, zdm = (double)acc[DQ_ZDM] / Ns;
	unsigned long long q2[2] = { 0, 0 };
	double t[3] = { 0, 0, 0 };	// GLV, ZDV, ZDE
	for (int lv = 0; lv < L1; lv++)
	{
		unsigned long long ri = 0;
		for (int d = c.tid; d < Nd; d += c.nthr)
		{
			const unsigned long long n = P[(size_t)lv * Nd + d];
			if (!n)
				continue;
			ri += n;
			atomicAdd (Cd + d, (uint32_t)n);
			const double di = (double)lv - glm, dj = (double)d - zdm, nd = (double)n;
			t[0] += nd * di * di;						// :468-471
			t[1] += nd * dj * dj;						// :474-476
			t[2] += nd * log2 (nd / Ns + EPS_ZONES);			// :446
		}
		unsigned long long rr[1] = { ri };
		cta_sum_u64<1> (c, rr);
		if (c.tid == 0)
			q2[0] += rr[0] * rr[0];						// :423
	}
	__syncthreads();
	for (int d = c.tid; d < Nd; d += c.nthr)
		q2[1] += (unsigned long long)Cd[d] * Cd[d];				// :404
	cta_sum_u64<2> (c, q2);
	cta_sum<3> (c, t);
	if (c.tid == 0)
	{
		const double zdnu = (double)q2[1] / Ns, glnu = (double)q2[0] / Ns;
		out[Z_SDE] = s[DA_SDE] / Ns;
		out[Z_LDE] = (double)acc[DQ_LDE] / Ns;
		out[Z_LGLZE] = s[DA_LGL] / Ns;
		out[Z_HGLZE] = (double)acc[DQ_HGL] / Ns;
		out[Z_SDLGLE] = s[DA_SDL] / Ns;
		out[Z_SDHGLE] = s[DA_SDH] / Ns;
		out[Z_LDLGLE] = s[DA_LDL] / Ns;
		out[Z_LDHGLE] = s[DA_LDH] / Ns;
		out[Z_GLNU] = glnu;
		out[Z_GLNUN] = glnu / Ns;					// :428
		out[Z_ZDNU] = zdnu;
		out[Z_ZDNUN] = zdnu / Ns;					// :411
		out[Z_ZP] = Ns / Nv;						// :455
		out[Z_GLM] = glm;
		out[Z_GLV] = t[0] / Ns;
		out[Z_ZDM] = zdm;
		out[Z_ZDV] = t[1] / Ns;
		out[Z_ZDE] = -(t[2] / Ns);
	}
}


// ------------------------------------------------------------------------------------------------
// GLDM / NGLDM / NGTDM: one 3x3 stencil pass over the global tile, matrices privatised in shared memory per CTA
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int big_nbr_same_level (const uint32_t* __restrict__ cell, int w, int h, int x, int y, uint32_t lv, uint32_t flag)
{
	int n = 0;
	for (int dy = -1; dy <= 1; dy++)
	{
		const int yy = y + dy;
		if (yy < 0 || yy >= h)
			continue;
		for (int dx = -1; dx <= 1; dx++)
		{
			const int xx = x + dx;
			if ((!dx && !dy) || xx < 0 || xx >= w)
				continue;
			const uint32_t q = cell[(unsigned long long)yy * w + xx];
			if ((q & flag) && (q & CELL_LEVEL) == lv)
				n++;
		}
	}
	return n;
}

// GLDM (ngldm = 0): centres / neighbours with CELL_NZ, column j = 1 + matches in 1..9, P[L1][10]
// NGLDM (ngldm = 1): centres / neighbours with CELL_IN, column = matches in 0..8, P[L1][10] (column 9 unused)
__global__ void __launch_bounds__(BG_T) k_big_dep_build (const uint32_t* __restrict__ cell, int w, int h, int L1, int ngldm, int use_sm, uint32_t* P)
{
	extern __shared__ __align__(16) uint32_t sm[];
	const unsigned long long ncell = (unsigned long long)w * (unsigned long long)h;
	const size_t nw = (size_t)L1 * 10;
	uint32_t* T = use_sm ? sm : P;
	if (use_sm)
	{
		for (size_t k = threadIdx.x; k < nw; k += blockDim.x)
			sm[k] = 0;
		__syncthreads();
	}
	const uint32_t flag = ngldm ? CELL_IN : CELL_NZ;
	CellIter it;
	for (it.init (ncell, w); it.valid(); it.next (w))
	{
		const uint32_t cc = cell[it.i];
		if (!(cc & flag))
			continue;
		const uint32_t lv = cc & CELL_LEVEL;
		const int nm = big_nbr_same_level (cell, w, h, it.col, it.row, lv, flag);
		atomicAdd (T + (size_t)lv * 10 + (ngldm ? nm : nm + 1), 1u);
	}
	if (use_sm)
	{
		__syncthreads();
		for (size_t k = threadIdx.x; k < nw; k += blockDim.x)
		{
			const uint32_t v = sm[k];
			if (v)
				atomicAdd (P + k, v);
		}
	}
}

// GLDM features from the dense P[L1][10] (gldm.cpp:295-554; nbr_body.h gldm_body)
__global__ void __launch_bounds__(BG_T) k_big_gldm_reduce (const uint32_t* __restrict__ P, int L1, Params prm, double* out)
{
	__shared__ double red[RED_MAX * NW_MAX];
	__shared__ unsigned long long Sj[10];
	const Cta c = big_cta (red);
	if (c.tid < 10)
		Sj[c.tid] = 0;
	__syncthreads();
	enum { Q_NZ, Q_LDE, Q_MUI, Q_MUJ, Q_HGL, Q_SI2, Q_N };
	unsigned long long q[Q_N];
	for (int k = 0; k < Q_N; k++)
		q[k] = 0;
	enum { A_SDE, A_LGL, A_SDL, A_SDH, A_LDL, A_LDH, A_N };
	double s[A_N];
	for (int k = 0; k < A_N; k++)
		s[k] = 0;
	for (int lv = c.tid; lv < L1; lv += c.nthr)
	{
		unsigned long long si = 0;
		const double in = (double)lv;
		for (int j = 1; j <= 9; j++)
		{
			const unsigned long long n = P[(size_t)lv * 10 + j];
			if (!n)
				continue;
			si += n;
			atomicAdd (&Sj[j], n);
			const double nd = (double)n;
			q[Q_NZ] += n;
			q[Q_LDE] += n * (unsigned long long)(j * j);
			q[Q_MUJ] += n * (unsigned long long)j;
			q[Q_MUI] += n * lv;
			q[Q_HGL] += n * (unsigned long long)lv * (unsigned long long)lv;
			s[A_SDE] += nd / ((double)j * (double)j);
			s[A_LGL] += nd / (in * in);
			s[A_SDL] += nd / (double)(in * in * j * j);
			s[A_SDH] += nd * (in * in) / (double)(j * j);
			s[A_LDL] += nd * (double)(j * j) / (in * in);
			s[A_LDH] += nd * (double)(in * in * j * j);
		}
		q[Q_SI2] += si * si;
	}
	__syncthreads();
	cta_sum_u64<Q_N> (c, q);
	cta_sum<A_N> (c, s);
	if (q[Q_NZ] == 0)	// gldm.cpp:213-231
	{
		for (int i = c.tid; i < NCOL_GLDM; i += c.nthr)
			out[i] = prm.softnan;
		return;
	}
	const double Nz = (double)q[Q_NZ];
	const double mui = (double)q[Q_MUI] / Nz, muj = (double)q[Q_MUJ] / Nz;
	double t[3] = { 0, 0, 0 };	// GLV, DV, DE
	for (int lv = c.tid; lv < L1; lv += c.nthr)
		for (int j = 1; j <= 9; j++)
		{
			const unsigned long long n = P[(size_t)lv * 10 + j];
			if (!n)
				continue;
			const double nd = (double)n, di = (double)lv - mui, dj = (double)j - muj;
			t[0] += nd * di * di;
			t[1] += nd * dj * dj;
			t[2] += nd * fast_log10 (dadd (nd / Nz, EPS_ZONES));
		}
	cta_sum<3> (c, t);
	if (c.tid == 0)
	{
		unsigned long long sj2 = 0;
		for (int j = 1; j <= 9; j++)
			sj2 += Sj[j] * Sj[j];
		out[D_SDE] = s[A_SDE] / Nz;
		out[D_LDE] = (double)q[Q_LDE] / Nz;
		out[D_GLN] = (double)q[Q_SI2] / Nz;
		out[D_DN] = (double)sj2 / Nz;
		out[D_DNN] = (double)sj2 / (Nz * Nz);
		out[D_GLV] = t[0] / Nz;
		out[D_DV] = t[1] / Nz;
		out[D_DE] = -(t[2] / Nz) * (1.0 / LOG10_2);
		out[D_LGLE] = s[A_LGL] / Nz;
		out[D_HGLE] = (double)q[Q_HGL] / Nz;
		out[D_SDLGLE] = s[A_SDL] / Nz;
		out[D_SDHGLE] = s[A_SDH] / Nz;
		out[D_LDLGLE] = s[A_LDL] / Nz;
		out[D_LDHGLE] = s[A_LDH] / Nz;
	}
}

// NGLDM features from the dense P[L1][10] (ngldm.cpp:267-355; nbr_body.h ngldm_body)
__global__ void __launch_bounds__(BG_T) k_big_ngldm_reduce (const uint32_t* __restrict__ P, int L1, Params prm, double* out)
{
	__shared__ double red[RED_MAX * NW_MAX];
	__shared__ unsigned long long Sr[10];
	const Cta c = big_cta (red);
	if (c.tid < 10)
		Sr[c.tid] = 0;
	__syncthreads();
	enum { Q_NS, Q_G, Q_DC, Q_SG2, Q_P2, Q_N };
	unsigned long long q[Q_N];
	for (int k = 0; k < Q_N; k++)
		q[k] = 0;
	enum { A_LDE, A_HDE, A_LGLCE, A_HGLCE, A_LDLGLE, A_LDHGLE, A_HDLGLE, A_HDHGLE, A_N };
	double s[A_N];
	for (int k = 0; k < A_N; k++)
		s[k] = 0;
	for (int lv = c.tid; lv < L1; lv += c.nthr)
	{
		unsigned long long sg = 0;
		const double g = (double)lv;
		for (int nm = 0; nm < 9; nm++)
		{
			const unsigned long long n = P[(size_t)lv * 10 + nm];
			if (!n)
				continue;
			sg += n;
			atomicAdd (&Sr[nm], n);
			const double nd = (double)n, dc = (double)(nm + 1);		// ngldm.cpp:283-305
			q[Q_NS] += n;
			q[Q_G] += n * lv;
			q[Q_DC] += n * (unsigned long long)(nm + 1);
			q[Q_P2] += n * n;
			s[A_LDE] += nd / dc / dc;
			s[A_HDE] += nd * dc * dc;
			if (g != 0.0)
			{
				s[A_LGLCE] += nd / g / g;
				s[A_LDLGLE] += nd / dc / dc / g / g;
				s[A_HDLGLE] += nd * dc * dc / g / g;
			}
			s[A_HGLCE] += nd * g * g;
			s[A_LDHGLE] += nd * g * g / dc / dc;
			s[A_HDHGLE] += nd * dc * dc * g * g;
		}
		q[Q_SG2] += sg * sg;
	}
	__syncthreads();
	cta_sum_u64<Q_N> (c, q);
	cta_sum<A_N> (c, s);
	const double Ns = (double)q[Q_NS];
	const double glm = (double)q[Q_G] / Ns, dcm = (double)q[Q_DC] / Ns;
	double t[3] = { 0, 0, 0 };
	for (int lv = c.tid; lv < L1; lv += c.nthr)
		for (int nm = 0; nm < 9; nm++)
		{
			const unsigned long long n = P[(size_t)lv * 10 + nm];
			if (!n)
				continue;
			const double nd = (double)n, dg = (double)lv - glm, dd = (double)(nm + 1) - dcm;
			t[0] += nd * dg * dg;						// :330
			t[1] += nd * dd * dd;						// :337
			t[2] += nd * log (nd / Ns);					// :302-303
		}
	cta_sum<3> (c, t);
	if (c.tid == 0)
	{
		unsigned long long sr2 = 0;
		for (int k = 0; k < 9; k++)
			sr2 += Sr[k] * Sr[k];
		out[N_LDE] = s[A_LDE] / Ns;
		out[N_HDE] = s[A_HDE] / Ns;
		out[N_LGLCE] = s[A_LGLCE] / Ns;
		out[N_HGLCE] = s[A_HGLCE] / Ns;
		out[N_LDLGLE] = s[A_LDLGLE] / Ns;
		out[N_LDHGLE] = s[A_LDHGLE] / Ns;
		out[N_HDLGLE] = s[A_HDLGLE] / Ns;
		out[N_HDHGLE] = s[A_HDHGLE] / Ns;
		out[N_GLNU] = (double)q[Q_SG2] / Ns;
		out[N_GLNUN] = (double)q[Q_SG2] / (Ns * Ns);
		out[N_DCNU] = (double)sr2 / Ns;
		out[N_DCNUN] = (double)sr2 / (Ns * Ns);
		out[N_DCP] = 1.0;						// :354
		out[N_GLM] = glm;
		out[N_GLV] = t[0] / Ns;
		out[N_DCM] = dcm;
		out[N_DCV] = t[1] / Ns;
		out[N_DCENT] = -(t[2] / Ns) / log (2.0);
		out[N_DCENE] = (double)q[Q_P2] / (Ns * Ns);
	}
}

// NGTDM build: N[L1] (u32), SI[L1][8] exact integer |level*nd - sum| (u64), PR[L1] presence; Nvc in acc[0]
__global__ void __launch_bounds__(BG_T) k_big_ngtdm_build (const uint32_t* __restrict__ cell, int w, int h, int L1, int use_sm,
	uint32_t* Nn, unsigned long long* SI, uint32_t* PR, unsigned long long* acc)
{
	extern __shared__ __align__(16) unsigned long long sm64[];	// SI[L1][8] (u64) then N[L1], PR[L1] (u32)
	__shared__ double red[RED_MAX * NW_MAX];
	const Cta c = big_cta (red);
	const unsigned long long ncell = (unsigned long long)w * (unsigned long long)h;
	unsigned long long* sSI = use_sm ? sm64 : SI;
	uint32_t* sN = use_sm ? (uint32_t*)(sm64 + (size_t)L1 * 8) : Nn;
	uint32_t* sPR = use_sm ? sN + L1 : PR;
	if (use_sm)
	{
		for (size_t k = threadIdx.x; k < (size_t)L1 * 8; k += blockDim.x)
			sSI[k] = 0;
		for (int k = threadIdx.x; k < 2 * L1; k += blockDim.x)
			sN[k] = 0;
		__syncthreads();
	}
	unsigned long long q[1] = { 0 };
	CellIter it;
	for (it.init (ncell, w); it.valid(); it.next (w))
	{
		const uint32_t lv = cell[it.i] & CELL_LEVEL;
		if (!lv)
			continue;
		sPR[lv] = 1;
		int nd = 0;
		unsigned long long sum = 0;
		for (int dy = -1; dy <= 1; dy++)
		{
			const int yy = it.row + dy;
			if (yy < 0 || yy >= h)
				continue;
			for (int dx = -1; dx <= 1; dx++)
			{
				const int xx = it.col + dx;
				if ((!dx && !dy) || xx < 0 || xx >= w)
					continue;
				const uint32_t qv = cell[(unsigned long long)yy * w + xx] & CELL_LEVEL;
				if (qv)
				{
					sum += qv;
					nd++;
				}
			}
		}
		if (nd > 0)	// ngtdm.cpp:83-146
		{
			const long long df = (long long)lv * nd - (long long)sum;
			atomicAdd (sN + lv, 1u);
			atomicAdd (sSI + (size_t)lv * 8 + (nd - 1), (unsigned long long)(df < 0 ? -df : df));
			q[0]++;
		}
	}
	if (use_sm)
	{
		__syncthreads();
		for (size_t k = threadIdx.x; k < (size_t)L1 * 8; k += blockDim.x)
			if (sSI[k])
				atomicAdd (SI + k, sSI[k]);
		for (int k = threadIdx.x; k < L1; k += blockDim.x)
		{
			if (sN[k])
				atomicAdd (Nn + k, sN[k]);
			if (sPR[k])
				PR[k] = 1;
		}
	}
	big_add_acc<1> (c, q, acc);
}

// NGTDM features (ngtdm.cpp:228-347; nbr_body.h ngtdm_reduce): levels compacted by one warp, pair loops over present levels.
// work: lst (u32[L1]) + Pl, Sl (double[L1]) in global scratch.
__global__ void __launch_bounds__(BG_T) k_big_ngtdm_reduce (const uint32_t* __restrict__ Nn, const unsigned long long* __restrict__ SI, const uint32_t* __restrict__ PR,
	int L1, const unsigned long long* acc, int ibsi, Params prm, uint32_t* lst, double* Pl, double* Sl, double* out)
{
	__shared__ double red[RED_MAX * NW_MAX];
	__shared__ int sh_cnt[2];
	const Cta c = big_cta (red);
	const double Nvc = (double)acc[0], Nvp = Nvc;
	if (c.tid < 32)
	{
		int Ngp = 0, nl = 0;
		for (int base = 0; base < L1; base += 32)
		{
			const int lv = base + c.tid;
			uint32_t n = 0, pr = 0;
			double sv = 0;
			if (lv < L1)
			{
				n = Nn[lv];
				pr = PR[lv];
				for (int k = 0; k < 8; k++)
				{
					const unsigned long long si = SI[(size_t)lv * 8 + k];
					if (si)
						sv += (double)si / (double)(k + 1);
				}
			}
			const bool keep = n != 0 && lv >= 1;
			const unsigned bk = __ballot_sync (0xffffffffu, keep), bp = __ballot_sync (0xffffffffu, pr != 0);
			if (keep)
			{
				const int pos = nl + __popc (bk & ((1u << c.tid) - 1u));
				lst[pos] = (uint32_t)lv;
				Pl[pos] = (double)n / Nvc;
				Sl[pos] = sv;
			}
			nl += __popc (bk);
			Ngp += __popc (bp);
		}
		if (c.tid == 0)
		{
			sh_cnt[0] = Ngp;
			sh_cnt[1] = nl;
		}
	}
	__syncthreads();
	const int Ngp = sh_cnt[0], nl = sh_cnt[1];
	const int Isize = ibsi ? (Ngp ? L1 : 0) : Ngp;		// |I| (ngtdm.cpp:57-67)
	if (Isize < 2)
	{
		for (int i = c.tid; i < NCOL_NGTDM; i += c.nthr)
			out[i] = prm.softnan;
		return;
	}
	enum { A_PS, A_S, A_CON, A_BUSY, A_CPLX, A_STR, A_N };
	double s[A_N];
	for (int k = 0; k < A_N; k++)
		s[k] = 0;
	for (int k = c.tid; k < nl; k += c.nthr)
	{
		s[A_PS] += Pl[k] * Sl[k];
		s[A_S] += Sl[k];
	}
	for (int a = 0; a < nl; a++)
	{
		const double pi = Pl[a], iv = (double)lst[a], sa = Sl[a];
		for (int b = c.tid; b < nl; b += c.nthr)
		{
			const double pj = Pl[b], jv = (double)lst[b];
			s[A_CON] += pi * pj * (iv - jv) * (iv - jv);
			s[A_BUSY] += fabs (pi * iv - pj * jv);
			s[A_CPLX] += fabs (iv - jv) * (pi * sa + pj * Sl[b]) / (pi + pj);
			s[A_STR] += (pi + pj) * (iv - jv) * (iv - jv);
		}
	}
	cta_sum<A_N> (c, s);
	if (c.tid == 0)
	{
		out[T_COARSENESS] = 1.0 / s[A_PS];
		const int ngp2 = Ngp > 1 ? Ngp * (Ngp - 1) : Ngp;
		out[T_CONTRAST] = (s[A_CON] / (double)ngp2) * (s[A_S] / Nvc);
		out[T_BUSYNESS] = (Ngp == 1 || s[A_BUSY] == 0) ? 0.0 : s[A_PS] / s[A_BUSY];
		out[T_COMPLEXITY] = s[A_CPLX] / Nvp;
		out[T_STRENGTH] = s[A_STR] / s[A_S];
	}
}


// ------------------------------------------------------------------------------------------------
// First-order statistics of a giant ROI: exact VALUE histogram (imax - imin + 1 integer counters) built straight from
// the slide, then every feature of intensity_body.h derived from the histogram by one CTA -- no pixel cloud, no sort
// (median / mode / robust means are rank queries on the histogram).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BG_T) k_big_int_hist (ImageView im, RoiRec r, uint32_t* hist)
{
	const unsigned long long ncell = (unsigned long long)r.w * (unsigned long long)r.h;
	CellIter it;
	for (it.init (ncell, r.w); it.valid(); it.next (r.w))
	{
		const size_t g = (size_t)(r.y0 + it.row) * im.W + (size_t)(r.x0 + it.col);
		if (im.label[g] == r.label)
			atomicAdd (hist + (load_inten (im, g) - r.imin), 1u);
	}
}

// value at rank k (0-based) of the sorted cloud; every thread gets the answer
__device__ uint32_t big_hist_select (const Cta& c, const uint32_t* __restrict__ hist, uint32_t nbins, unsigned long long k, unsigned long long* sh_tot, uint32_t* sh_ans)
{
	const uint32_t chunk = (nbins + c.nthr - 1) / c.nthr;
	const uint32_t b0 = min (nbins, (uint32_t)c.tid * chunk), b1 = min (nbins, b0 + chunk);
	unsigned long long tot = 0;
	for (uint32_t b = b0; b < b1; b++)
		tot += hist[b];
	sh_tot[c.tid] = tot;
	__syncthreads();
	if (c.tid == 0)
	{
		unsigned long long run = 0;
		for (int t = 0; t < c.nthr; t++)
		{
			const unsigned long long x = sh_tot[t];
			sh_tot[t] = run;
			run += x;
		}
	}
	__syncthreads();
	const unsigned long long ex = sh_tot[c.tid];
	if (k >= ex && k < ex + tot)
	{
		unsigned long long run = ex;
		for (uint32_t b = b0; b < b1; b++)
		{
			run += hist[b];
			if (k < run)
			{
				*sh_ans = b;
				break;
			}
		}
	}
	__syncthreads();
	const uint32_t ans = *sh_ans;
	__syncthreads();
	return ans;
}

__global__ void __launch_bounds__(BG_T) k_big_int_reduce (const uint32_t* __restrict__ hist, RoiRec r, Params prm, double* out)
{
	extern __shared__ __align__(16) uint32_t sm[];		// b100[101] + bc[nb+1]
	__shared__ double red[RED_MAX * NW_MAX];
	__shared__ unsigned long long sh_tot[BG_T];
	__shared__ uint32_t sh_ans;
	__shared__ double shp[6];
	const Cta c = big_cta (red);
	const int nb = prm.greydepth > 0 ? prm.greydepth : -prm.greydepth;
	uint32_t* b100 = sm;
	uint32_t* bc = b100 + 101;
	const uint32_t mn = r.imin, mx = r.imax, range = mx - mn, nbins = range + 1;
	const unsigned long long n = r.area;
	const double nd = (double)n;
	for (int i = c.tid; i < NCOL_INT; i += c.nthr)
		out[i] = 0.0;
	for (int i = c.tid; i < 101 + nb + 1; i += c.nthr)
		sm[i] = 0;
	__syncthreads();

	// ---- pass 1: exact raw sums (intensity.cpp:86-98)
	unsigned long long s12[2] = { 0, 0 };
	for (uint32_t b = c.tid; b < nbins; b += c.nthr)
	{
		const unsigned long long cnt = hist[b];
		if (!cnt)
			continue;
		const uint32_t v = mn + b;
		s12[0] += cnt * v;
		s12[1] += cnt * (unsigned long long)(uint32_t)(v * v);
	}
	cta_sum_u64<2> (c, s12);
	const double mean = (double)s12[0] / nd;
	const double energy = (double)s12[1];

	// ---- pass 2: central sums + the two binned histograms (intensity.cpp:101-110,169-182; histogram.h:50-77)
	const double binW100 = ddiv ((double)range, 100.0);
	enum { C_MAD, C_M2, C_M3, C_M4, C_P5, C_P6, C_N };
	double cs[C_N] = { 0, 0, 0, 0, 0, 0 };
	const bool blank = (mn == 0 && mx == 0);
	for (uint32_t b = c.tid; b < nbins; b += c.nthr)
	{
		const uint32_t cnt = hist[b];
		if (!cnt)
			continue;
		const uint32_t v = mn + b;
		const double cd = (double)cnt, d = (double)v - mean, d2 = d * d;
		cs[C_MAD] += cd * fabs (d);
		cs[C_M2] += cd * d2;
		cs[C_M3] += cd * (d2 * d);
		cs[C_M4] += cd * (d2 * d2);
		cs[C_P5] += cd * (d2 * d2 * d);
		cs[C_P6] += cd * (d2 * d2 * d2);
		const double ri = ddiv ((double)(uint32_t)(v - mn), binW100);
		const int idx = (ri == ri) ? (int)ri : 0;
		atomicAdd (b100 + idx, cnt);
		atomicAdd (bc + to_grayscale (v, mn, range, (uint32_t)nb, false), cnt);
	}
	__syncthreads();
	cta_sum<C_N> (c, cs);
	const double var = cs[C_M2];
	const double variance = n > 1 ? var / (nd - 1.0) : 0.0;
	const double variance_b = n > 1 ? var / nd : 0.0;
	const double sd = sqrt (variance), sd_b = sqrt (variance_b);
	if (c.tid == 0)
	{
		out[I_MIN] = (double)mn;
		out[I_MAX] = (double)mx;
		out[I_RANGE] = (double)mx - (double)mn;
		out[I_COVERED_IMAGE_INTENSITY_RANGE] = prm.covered_den != 0.0 ? (double)(uint32_t)(mx - mn) / prm.covered_den : -0.0;
		out[I_MEAN] = mean;
		out[I_ENERGY] = energy;
		out[I_ROOT_MEAN_SQUARED] = sqrt (energy / nd);
		out[I_INTEGRATED_INTENSITY] = (double)s12[0];
		out[I_MEAN_ABSOLUTE_DEVIATION] = cs[C_MAD] / nd;
		out[I_VARIANCE] = variance;
		out[I_VARIANCE_BIASED] = variance_b;
		out[I_STANDARD_DEVIATION] = sd;
		out[I_STANDARD_DEVIATION_BIASED] = sd_b;
		out[I_COV] = sd / mean;
		out[I_STANDARD_ERROR] = sd / sqrt (nd);
	}
	double* hist_out = prm.hist_col ? out + (prm.hist_col - COL_INT) : nullptr;	// HISTOGRAM_BIN_* (see intensity_body.h)
	if (blank)	// intensity.cpp:121-122
	{
		if (hist_out)
			for (int k = c.tid; k < nb; k += c.nthr)
				hist_out[k] = 0.0;
		return;
	}
	if (c.tid == 0)
	{
		b100[99] += b100[100];
		b100[100] = 0;
		bc[nb - 1] += bc[nb];
		bc[nb] = 0;
	}
	__syncthreads();
	if (hist_out)
		for (int k = c.tid; k < nb; k += c.nthr)
			hist_out[k] = (double)bc[k];
	{
		const double qq[6] = { 0.01, 0.1, 0.25, 0.75, 0.9, 0.99 };
		if (c.tid < 6)
			shp[c.tid] = hist_percentile (b100, dmul (nd, qq[c.tid]), binW100, mn);
	}
	double eu[2] = { 0, 0 };
	for (int k = c.tid; k <= nb; k += c.nthr)
	{
		const double p = (double)bc[k] / nd;
		eu[0] += p * log2 (p + 2.2e-16);
		eu[1] += p * p;
	}
	__syncthreads();
	cta_sum<2> (c, eu);
	const double p01 = shp[0], p10 = shp[1], p25 = shp[2], p75 = shp[3], p90 = shp[4], p99 = shp[5];

	// ---- median (histogram.h:268-287) and mode (:289-309) as rank / arg-max queries
	const uint32_t v_hi = mn + big_hist_select (c, hist, nbins, n / 2, sh_tot, &sh_ans);
	double median;
	if (n & 1)
		median = (double)v_hi;
	else
	{
		const uint32_t v_lo = mn + big_hist_select (c, hist, nbins, n / 2 - 1, sh_tot, &sh_ans);
		median = (double)(uint32_t)(v_hi + v_lo) / 2.0;
	}
	unsigned long long best = 0;
	for (uint32_t b = c.tid; b < nbins; b += c.nthr)
	{
		const unsigned long long cnt = hist[b];
		if (cnt)
		{
			const unsigned long long key = (cnt << 32) | (unsigned long long)(~(mn + b));
			best = key > best ? key : best;
		}
	}
	best = cta_max_u64 (c, best);
	const double mode = (double)(uint32_t)(~(uint32_t)(best & 0xFFFFFFFFull));

	// ---- robust mean (intensity.cpp:139-149), median absolute deviation (:156-159)
	unsigned long long rb[2] = { 0, 0 };
	double md[1] = { 0 };
	for (uint32_t b = c.tid; b < nbins; b += c.nthr)
	{
		const unsigned long long cnt = hist[b];
		if (!cnt)
			continue;
		const double a = (double)(mn + b);
		if (a >= p10 && a <= p90)
		{
			rb[0] += cnt * (mn + b);
			rb[1] += cnt;
		}
		md[0] += (double)cnt * fabs (a - median);
	}
	cta_sum_u64<2> (c, rb);
	cta_sum<1> (c, md);
	const double robust_mean = rb[1] ? (double)rb[0] / (double)rb[1] : 0.0;
	double rm[1] = { 0 };
	if (rb[1])
		for (uint32_t b = c.tid; b < nbins; b += c.nthr)
		{
			const unsigned long long cnt = hist[b];
			const double a = (double)(mn + b);
			if (cnt && a >= p10 && a <= p90)
				rm[0] += (double)cnt * fabs (a - robust_mean);
		}
	cta_sum<1> (c, rm);
	const double rmad = rb[1] ? rm[0] / (double)rb[1] : 0.0;
	if (c.tid == 0)
	{
		out[I_MEDIAN] = median;
		out[I_MODE] = mode;
		out[I_P01] = p01;
		out[I_P10] = p10;
		out[I_P25] = p25;
		out[I_P75] = p75;
		out[I_P90] = p90;
		out[I_P99] = p99;
		out[I_QCOD] = (p75 - p25) / (p75 + p25);
		out[I_INTERQUARTILE_RANGE] = p75 - p25;
		out[I_ROBUST_MEAN] = robust_mean;
		out[I_ROBUST_MEAN_ABSOLUTE_DEVIATION] = rmad;
		out[I_ENTROPY] = -eu[0];
		out[I_UNIFORMITY] = eu[1];
		out[I_MEDIAN_ABSOLUTE_DEVIATION] = md[0] / nd;
		out[I_UNIFORMITY_PIU] = (1.0 - (double)(uint32_t)(mx - mn) / (double)(uint32_t)(mx + mn)) * 100.0;
		const double M2 = cs[C_M2], M3 = cs[C_M3], M4 = cs[C_M4];
		double skew = 0, kurt = 0, ekurt = 0;
		if (M2 != 0.0)
		{
			skew = n > 3 ? (sqrt (nd) * M3) / pow (M2, 1.5) : 0.0;
			kurt = n > 4 ? (nd * M4) / (M2 * M2) : 0.0;
			ekurt = n > 4 ? (nd * M4) / (M2 * M2) - 3.0 : 0.0;
		}
		out[I_SKEWNESS] = skew;
		out[I_KURTOSIS] = kurt;
		out[I_EXCESS_KURTOSIS] = ekurt;
		const double den5 = nd * pow (sd, 5.0), den6 = nd * pow (sd, 6.0);
		out[I_HYPERSKEWNESS] = den5 == 0.0 ? 0.0 : cs[C_P5] / den5;
		out[I_HYPERFLATNESS] = den6 == 0.0 ? 0.0 : cs[C_P6] / den6;
	}
}

} // namespace nyx
