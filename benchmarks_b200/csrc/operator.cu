// operator.cu -- C ABI section 4: geometry, the L-vector operator object (vmult & friends).
//
// Host-side mirror of Portable::LaplaceOperator (CEED_bp/include/portable_laplace_operator.h:17-96):
//   ctor            -> b200fe_op_create           (:98-122)
//   vmult           -> b200fe_op_vmult            (:124-172)
//   vmult_dummy     -> b200fe_op_vmult_dummy      (:175-235)
//   compute_G_tensors -> b200fe_geometry_from_nodes (:239-302, math of
//                        bakeoff_problems_dealii/include/portable_laplace_operator.h:227-258)
//   compute_diagonal  -> b200fe_op_diagonal       (bp5_kokkos/benchmark.cc:218-251)
//   compute_rhs       -> b200fe_op_rhs_one        (CEED_bp/src/bp3.cc:184-239)
#include <cmath>
#include <cstring>
#include <memory>

#include "halo.h"
#include "operator.h"
#include "mesh_common.h"

namespace b200fe {

unsigned long long g_launch_count = 0;

// ---------------------------------------------------------------------------------------------
// setup kernels (not on the timed path): simple, runtime-sized
// ---------------------------------------------------------------------------------------------
namespace {

struct GeoMats {          // geometry basis at the quadrature points, ng <= 9, nq <= 10
    double V[10 * 9];     // V[q*ng+a]
    double dV[10 * 9];
    double w[10];
};

// mapping support points of a (possibly smoothly deformed) box mesh: nodes[cell][d][c][b][a]
// cell_xyz[cell][stride]: stride 3 = (x,y,z) on the uniform mesh; stride 4 = (level,x,y,z), a level-1 cell is half the size
__global__ void box_nodes_kernel(uint32_t n_cells, int ng, const int32_t *__restrict__ cell_xyz, int stride,
                                 double p1x, double p1y, double p1z, double hx, double hy, double hz,
                                 const double *__restrict__ t /*[ng] GLL on [0,1]*/, int deform_kind,
                                 double amp, double freq, double *__restrict__ nodes)
{
    const int ng3 = ng * ng * ng;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint64_t)n_cells * ng3) return;
    const uint32_t cell = (uint32_t)(i / ng3);
    const int n = (int)(i % ng3), a = n % ng, b = (n / ng) % ng, c = n / (ng * ng);
    const int32_t *cx = cell_xyz + (size_t)cell * stride + (stride - 3);
    const double sc = (stride == 4 && cell_xyz[(size_t)cell * 4] == 1) ? 0.5 : 1.0;
    double x = p1x + (cx[0] + t[a]) * (hx * sc);
    double y = p1y + (cx[1] + t[b]) * (hy * sc);
    double z = p1z + (cx[2] + t[c]) * (hz * sc);
    if (deform_kind == 1) {  // smooth volume-preserving-ish perturbation, same family as bk3_dealii/check_bk3.cc:50-52
        const double dx = amp * sin(freq * y), dy = amp * sin(freq * z), dz = amp * sin(freq * x);
        x += dx; y += dy; z += dz;
    }
    double *o = nodes + (size_t)cell * 3 * ng3 + n;
    o[0] = x; o[ng3] = y; o[2 * ng3] = z;
}

// one CTA per cell, one thread per quadrature point (looping): J = sum_nodes x_node grad phi_node
__global__ void geometry_kernel(const __grid_constant__ GeoMats gm, uint32_t n_cells, int ng, int nq,
                                const double *__restrict__ nodes, double *__restrict__ G,
                                double *__restrict__ JxW)
{
    extern __shared__ double sn[];  // 3*ng^3
    const int ng3 = ng * ng * ng, nq3 = nq * nq * nq;
    for (uint32_t cell = blockIdx.x; cell < n_cells; cell += gridDim.x) {
        __syncthreads();
        for (int i = threadIdx.x; i < 3 * ng3; i += blockDim.x) sn[i] = nodes[(size_t)cell * 3 * ng3 + i];
        __syncthreads();
        for (int pt = threadIdx.x; pt < nq3; pt += blockDim.x) {
            const int qx = pt % nq, qy = (pt / nq) % nq, qz = pt / (nq * nq);  // point index p*nq^2+q*nq+r, p<->z
            double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};  // J[a][b] = d x_a / d xi_b, xi = (x^,y^,z^)
            for (int c = 0; c < ng; ++c)
                for (int b = 0; b < ng; ++b) {
                    const double vz = gm.V[qz * ng + c], dz = gm.dV[qz * ng + c];
                    const double vy = gm.V[qy * ng + b], dy = gm.dV[qy * ng + b];
                    for (int a = 0; a < ng; ++a) {
                        const double vx = gm.V[qx * ng + a], dx = gm.dV[qx * ng + a];
                        const double g0 = dx * vy * vz, g1 = vx * dy * vz, g2 = vx * vy * dz;
                        const int n = a + ng * (b + ng * c);
#pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            const double xn = sn[d * ng3 + n];
                            J[d][0] = fma(xn, g0, J[d][0]);
                            J[d][1] = fma(xn, g1, J[d][1]);
                            J[d][2] = fma(xn, g2, J[d][2]);
                        }
                    }
                }
            const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) -
                               J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                               J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
            const double id = 1.0 / det;
            double K[3][3];  // K[b][a] = d xi_b / d x_a
            K[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) * id;
            K[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
            K[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
            K[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) * id;
            K[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
            K[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
            K[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) * id;
            K[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
            K[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
            const double jxw = det * gm.w[qx] * gm.w[qy] * gm.w[qz];
            if (JxW) JxW[(size_t)cell * nq3 + pt] = jxw;
            if (G) {
                // kernel directions (r,s,t) = (z^, y^, x^): rows of K in the order 2,1,0
                const int perm[3] = {2, 1, 0};
                int ci = 0;
                for (int a = 0; a < 3; ++a)
                    for (int b = a; b < 3; ++b, ++ci) {
                        const double *ka = K[perm[a]], *kb = K[perm[b]];
                        G[((size_t)cell * 6 + ci) * nq3 + pt] = jxw * (ka[0] * kb[0] + ka[1] * kb[1] + ka[2] * kb[2]);
                    }
            }
        }
    }
}

// affine cells: J from the vertex differences of the trilinear map (columns v1-v0, v2-v0, v4-v0)
__global__ void affine_geometry_kernel(uint32_t n_cells, const double *__restrict__ nodes /*[cell][3][2][2][2]*/,
                                       double *__restrict__ cellG)
{
    const uint32_t cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= n_cells) return;
    const double *x = nodes + (size_t)cell * 24;
    double J[3][3];
    for (int d = 0; d < 3; ++d) {
        const double *xd = x + d * 8;  // node index c*4 + b*2 + a
        J[d][0] = xd[1] - xd[0];
        J[d][1] = xd[2] - xd[0];
        J[d][2] = xd[4] - xd[0];
    }
    const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                       J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
    const double id = 1.0 / det;
    double K[3][3];
    K[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) * id;
    K[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
    K[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
    K[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) * id;
    K[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
    K[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
    K[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) * id;
    K[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
    K[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
    const int perm[3] = {2, 1, 0};  // (r,s,t) = (z^, y^, x^), as in geometry_kernel
    double *o = cellG + (size_t)cell * 8;
    int ci = 0;
    for (int a = 0; a < 3; ++a)
        for (int b = a; b < 3; ++b, ++ci) {
            const double *ka = K[perm[a]], *kb = K[perm[b]];
            o[ci] = det * (ka[0] * kb[0] + ka[1] * kb[1] + ka[2] * kb[2]);
        }
    o[6] = det;
    o[7] = 0.0;
}

// deal.II MatrixFree data -> G: inv_jacobian(q, cell, ref, real) and JxW(q, cell) as Kokkos LayoutLeft views (q fastest)
__global__ void g_from_inv_jacobian_kernel(uint32_t n_cells, int nq3, const double *__restrict__ K, const double *__restrict__ JxW,
                                           double *__restrict__ G)
{
    const size_t total = (size_t)n_cells * nq3, plane = total;  // plane: stride between (ref, real) entries
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        double k[3][3];  // k[ref][real]
#pragma unroll
        for (int real = 0; real < 3; ++real)
#pragma unroll
            for (int ref = 0; ref < 3; ++ref) k[ref][real] = K[i + plane * (ref + 3 * real)];
        const double jxw = JxW[i];
        const size_t cell = i / nq3, q = i - cell * nq3;
        double *g = G + cell * 6 * (size_t)nq3 + q;
        const int perm[3] = {2, 1, 0};  // kernel directions (r,s,t) = (z^, y^, x^), as in geometry_kernel
        int ci = 0;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = a; b < 3; ++b, ++ci) {
                const double *ka = k[perm[a]], *kb = k[perm[b]];
                g[(size_t)ci * nq3] = jxw * (ka[0] * kb[0] + ka[1] * kb[1] + ka[2] * kb[2]);
            }
    }
}

__global__ void copy_constrained_kernel(uint32_t n, const uint32_t *__restrict__ list, size_t stride,
                                        const double *__restrict__ src, double *__restrict__ dst,
                                        double *__restrict__ dot, const int *__restrict__ skip)
{
    if (skip != nullptr && *skip != 0) return;
    src += blockIdx.y * stride; dst += blockIdx.y * stride;  // blockIdx.y = component
    double s = 0.0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t c = list[i];
        const double v = src[c];
        dst[c] = v;
        s = fma(v, v, s);
    }
    if (dot != nullptr) {
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((threadIdx.x & 31) == 0 && s != 0.0) atomicAdd(dot, s);
    }
}

// mats = shape_values[nm*nq] | co_shape_gradients[nq*nq] | shape_gradients[nm*nq]   (deal.II layouts)
// diag_local[l] = sum_q sum_ab G_ab d_a phi_l d_b phi_l (+ JxW phi_l^2), scattered with atomics
__global__ void diagonal_kernel(uint32_t n_cells, int nm, int nq, int qop, const double *__restrict__ mats,
                                const double *__restrict__ G, const double *__restrict__ JxW,
                                const uint32_t *__restrict__ idx, double *__restrict__ diag, const uint8_t *__restrict__ skip_cell)
{
    const int nm3 = nm * nm * nm, nq3 = nq * nq * nq;
    const double *S = mats, *Sg = mats + nm * nq + nq * nq;
    for (uint32_t cell = blockIdx.x; cell < n_cells; cell += gridDim.x) {
        if (skip_cell != nullptr && skip_cell[cell]) continue;  // interface cells of a constrained operator: handled as C^T A_c C
        for (int l = threadIdx.x; l < nm3; l += blockDim.x) {
            const uint32_t id = idx[(size_t)cell * nm3 + l];
            if (id == kInvalidIndex) continue;
            const int k = l % nm, j = (l / nm) % nm, i = l / (nm * nm);
            double s = 0.0;
            for (int p = 0; p < nq; ++p)
                for (int q = 0; q < nq; ++q)
                    for (int r = 0; r < nq; ++r) {
                        const int pt = (p * nq + q) * nq + r;
                        const double bi = S[i * nq + p], bj = S[j * nq + q], bk = S[k * nq + r];
                        if (qop & QOP_LAPLACE) {
                            const double gr = Sg[i * nq + p] * bj * bk, gs = bi * Sg[j * nq + q] * bk, gt = bi * bj * Sg[k * nq + r];
                            const double *g = G + (size_t)cell * 6 * nq3 + pt;
                            s += g[0] * gr * gr + g[3 * nq3] * gs * gs + g[5 * nq3] * gt * gt +
                                 2.0 * (g[nq3] * gr * gs + g[2 * nq3] * gr * gt + g[4 * nq3] * gs * gt);
                        }
                        if (qop & QOP_MASS) {
                            const double v = bi * bj * bk;
                            s += JxW[(size_t)cell * nq3 + pt] * v * v;
                        }
                    }
            atomicAdd(diag + id, s);
        }
    }
}

// ---- diagonal of C^T A C on the cells that hold hanging DoFs (b200fe_op_diagonal with constraints) -------------------
// MatrixFreeTools::compute_diagonal semantics (bp5_kokkos/benchmark.cc:231): column b of the cell matrix from the cell
// kernel applied to the local unit vector e_b, the local constraint matrix C_c applied on both sides, diagonal kept.
// gather the geometric factors of the listed cells into a contiguous block: dst[i][n] = src[cells[i]][n]
__global__ void gather_cells_kernel(uint32_t n_list, const uint32_t *__restrict__ cells, size_t per_cell, const double *__restrict__ src,
                                    double *__restrict__ dst)
{
    const size_t total = (size_t)n_list * per_cell;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t c = i / per_cell;
        dst[i] = src[(size_t)cells[c] * per_cell + (i - c * per_cell)];
    }
}

// src = e_b in every listed cell (element-vector layout [cell][nm^3]; masked local DoFs stay 0), dst = 0
__global__ void unit_vectors_kernel(uint32_t n_list, int nm3, int b, const uint32_t *__restrict__ idx_if, double *__restrict__ src,
                                    double *__restrict__ dst)
{
    const size_t total = (size_t)n_list * nm3;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int l = (int)(i % nm3);
        src[i] = (l == b && idx_if[i] != kInvalidIndex) ? 1.0 : 0.0;
        dst[i] = 0.0;
    }
}

// one CTA per listed cell: t = C_c^T (A_c e_b) into shared memory (compact target ids), then diag[target] += C_c(b, target) t[target]
__global__ void diag_constrained_reduce_kernel(uint32_t n_list, int nm3, int b, const double *__restrict__ col, const uint32_t *__restrict__ ent_ptr,
                                               const uint16_t *__restrict__ ent_cid, const double *__restrict__ ent_w,
                                               const uint32_t *__restrict__ tgt_ptr, const uint32_t *__restrict__ tgt, double *__restrict__ diag)
{
    extern __shared__ double t_sh[];
    for (uint32_t c = blockIdx.x; c < n_list; c += gridDim.x) {
        const uint32_t rb = ent_ptr[(size_t)c * nm3 + b], re = ent_ptr[(size_t)c * nm3 + b + 1];
        if (rb == re) continue;  // masked local DoF: no column (uniform across the CTA)
        const uint32_t t0 = tgt_ptr[c], nt = tgt_ptr[c + 1] - t0;
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < nt; i += blockDim.x) t_sh[i] = 0.0;
        __syncthreads();
        for (int a = threadIdx.x; a < nm3; a += blockDim.x) {
            const double v = col[(size_t)c * nm3 + a];
            for (uint32_t e = ent_ptr[(size_t)c * nm3 + a]; e < ent_ptr[(size_t)c * nm3 + a + 1]; ++e) atomicAdd(&t_sh[ent_cid[e]], ent_w[e] * v);
        }
        __syncthreads();
        for (uint32_t e = rb + threadIdx.x; e < re; e += blockDim.x) atomicAdd(diag + tgt[t0 + ent_cid[e]], ent_w[e] * t_sh[ent_cid[e]]);
    }
}

// b[l] += sum_q JxW(q) phi_l(q)    (bp3.cc:208-224: f = 1, constrained rows dropped)
__global__ void rhs_one_kernel(uint32_t n_cells, int nm, int nq, const double *__restrict__ mats,
                               const double *__restrict__ JxW, const uint32_t *__restrict__ idx,
                               double *__restrict__ b)
{
    const int nm3 = nm * nm * nm, nq3 = nq * nq * nq;
    const double *S = mats;
    for (uint32_t cell = blockIdx.x; cell < n_cells; cell += gridDim.x)
        for (int l = threadIdx.x; l < nm3; l += blockDim.x) {
            const uint32_t id = idx[(size_t)cell * nm3 + l];
            if (id == kInvalidIndex) continue;
            const int k = l % nm, j = (l / nm) % nm, i = l / (nm * nm);
            double s = 0.0;
            for (int p = 0; p < nq; ++p)
                for (int q = 0; q < nq; ++q) {
                    const double bij = S[i * nq + p] * S[j * nq + q];
                    for (int r = 0; r < nq; ++r)
                        s = fma(JxW[(size_t)cell * nq3 + (p * nq + q) * nq + r], bij * S[k * nq + r], s);
                }
            atomicAdd(b + id, s);
        }
}

// Setup of KArgs::excl_interior.  Pass 1 over the index table: `seen` / `multi` = DoF referenced at least once / twice,
// `inter` = DoF referenced from an interior position of a cell (all three local indices in 1..nm-2).
__global__ void excl_scan_kernel(size_t n_entries, int nm, const uint32_t *__restrict__ idx, uint32_t *seen, uint32_t *multi, uint32_t *inter)
{
    const int nm3 = nm * nm * nm;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n_entries; t += (size_t)gridDim.x * blockDim.x) {
        const uint32_t id = idx[t];
        if (id == kInvalidIndex) continue;
        const int l = (int)(t % nm3), k = l % nm, j = (l / nm) % nm, i = l / (nm * nm);
        const uint32_t bit = 1u << (id & 31);
        if (atomicOr(seen + (id >> 5), bit) & bit) atomicOr(multi + (id >> 5), bit);
        if (i >= 1 && i <= nm - 2 && j >= 1 && j <= nm - 2 && k >= 1 && k <= nm - 2) atomicOr(inter + (id >> 5), bit);
    }
}
// Pass 2: the property holds iff no DoF is both interior somewhere and referenced twice
__global__ void excl_check_kernel(uint32_t n_words, const uint32_t *__restrict__ multi, const uint32_t *__restrict__ inter, int *violation)
{
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += gridDim.x * blockDim.x)
        if (multi[w] & inter[w]) *violation = 1;
}

// Index table of a box mesh on the device (the device twin of BoxMesh::expand_indices, mesh.cc): entry (cell, l) = local
// index of the first DoF of the entity local DoF l sits on + its index inside the entity; invalid on the Dirichlet boundary.
// ent_idx[l] = entity << 16 | index in entity.
__global__ void expand_indices_kernel(size_t n_entries, int nm, int p, int dirichlet, int64_t dx, int64_t dy, int64_t dz,
                                      const uint32_t *__restrict__ lbase, const int32_t *__restrict__ cell_xyz,
                                      const uint32_t *__restrict__ ent_idx, uint32_t *__restrict__ out)
{
    const int nm3 = nm * nm * nm;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n_entries; t += (size_t)gridDim.x * blockDim.x) {
        const size_t cell = t / nm3;
        const int l = (int)(t - cell * nm3);
        const uint32_t ei = ent_idx[l];
        uint32_t loc = lbase[cell * 27 + (ei >> 16)] + (ei & 0xFFFFu);
        if (dirichlet) {
            const int a = l % nm, b = (l / nm) % nm, c = l / (nm * nm);
            const int64_t X = (int64_t)cell_xyz[cell * 3] * p + a, Y = (int64_t)cell_xyz[cell * 3 + 1] * p + b, Z = (int64_t)cell_xyz[cell * 3 + 2] * p + c;
            if (X == 0 || Y == 0 || Z == 0 || X == dx || Y == dy || Z == dz) loc = kInvalidIndex;
        }
        out[t] = loc;
    }
}

// Axis-aligned cells: diagonal and int phi_i of the separable operator from the diagonals of the 1-D matrices.
struct CartVecs { double dk[9], dm[9], mv[9]; };  // diag K, diag M, m = B^T w (nm <= 9 entries each)
__global__ void cart_diagonal_kernel(const __grid_constant__ CartVecs v, size_t n_entries, int nm, int qop, const double *__restrict__ cellG,
                                     const uint32_t *__restrict__ idx, double *diag)
{
    const int nm3 = nm * nm * nm;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n_entries; t += (size_t)gridDim.x * blockDim.x) {
        const uint32_t id = idx[t];
        if (id == kInvalidIndex) continue;
        const size_t cell = t / nm3;
        const int l = (int)(t - cell * nm3), k = l % nm, j = (l / nm) % nm, i = l / (nm * nm);
        const double *g = cellG + cell * 8;
        double d = 0.0;
        if (qop & QOP_LAPLACE) d = g[0] * v.dk[i] * v.dm[j] * v.dm[k] + g[3] * v.dm[i] * v.dk[j] * v.dm[k] + g[5] * v.dm[i] * v.dm[j] * v.dk[k];
        if (qop & QOP_MASS) d += g[6] * v.dm[i] * v.dm[j] * v.dm[k];
        atomicAdd(diag + id, d);
    }
}
__global__ void cart_rhs_kernel(const __grid_constant__ CartVecs v, size_t n_entries, int nm, const double *__restrict__ cellG,
                                const uint32_t *__restrict__ idx, double *b)
{
    const int nm3 = nm * nm * nm;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n_entries; t += (size_t)gridDim.x * blockDim.x) {
        const uint32_t id = idx[t];
        if (id == kInvalidIndex) continue;
        const size_t cell = t / nm3;
        const int l = (int)(t - cell * nm3), k = l % nm, j = (l / nm) % nm, i = l / (nm * nm);
        atomicAdd(b + id, cellG[cell * 8 + 6] * v.mv[i] * v.mv[j] * v.mv[k]);
    }
}

// *flag = 1 unless every cell's constants are those of an axis-aligned box (vanishing rs, rt, st couplings)
__global__ void cartesian_check_kernel(uint32_t n_cells, const double *__restrict__ cellG, int *flag)
{
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < n_cells; c += gridDim.x * blockDim.x) {
        const double *g = cellG + (size_t)c * 8;
        const double diag = fabs(g[0]) + fabs(g[3]) + fabs(g[5]), off = fabs(g[1]) + fabs(g[2]) + fabs(g[4]);
        if (!(off <= 1e-14 * diag)) *flag = 1;
    }
}

__global__ void set_constrained_kernel(uint32_t n, const uint32_t *__restrict__ list, double value, double *__restrict__ v)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[list[i]] = value;
}

// hanging-node rows, one warp per row (rows hold 2 ... (p+1)^2 parents): v[h] = sum_k w[k] v[col[k]].
// Vector-valued problems: the row is applied to all components (component-blocked vectors, `stride` apart) while its
// weights / parent indices are read once (the CSR arrays are the only DRAM traffic of these kernels; r01h: 87 us per
// component for 130 k rows at p = 8 when launched per component).
constexpr int kCompChunk = 4;

__global__ void distribute_kernel(uint32_t n_rows, const uint32_t *__restrict__ hdof, const uint32_t *__restrict__ ptr,
                                  const uint32_t *__restrict__ col, const double *__restrict__ w, double *__restrict__ v,
                                  int ncomp, size_t stride, double *__restrict__ save, const int *__restrict__ skip)
{
    if (skip != nullptr && *skip != 0) return;
    const int lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rows; r += warps) {
        const uint32_t b = ptr[r], e = ptr[r + 1], h = hdof[r];
        for (int c0 = 0; c0 < ncomp; c0 += kCompChunk) {
            const int nc = min(kCompChunk, ncomp - c0);
            double s[kCompChunk] = {0.0, 0.0, 0.0, 0.0};
            for (uint32_t k = b + lane; k < e; k += 32) {
                const double wk = w[k];
                const double *vk = v + (size_t)c0 * stride + col[k];
#pragma unroll
                for (int c = 0; c < kCompChunk; ++c)
                    if (c < nc) s[c] = fma(wk, vk[c * stride], s[c]);
            }
#pragma unroll
            for (int c = 0; c < kCompChunk; ++c)
                for (int o = 16; o > 0; o >>= 1) s[c] += __shfl_xor_sync(0xffffffffu, s[c], o);
            if (lane == 0) {
#pragma unroll
                for (int c = 0; c < kCompChunk; ++c)
                    if (c < nc) {
                        double *vh = v + (size_t)(c0 + c) * stride + h;
                        if (save) save[(size_t)(c0 + c) * n_rows + r] = *vh;
                        *vh = s[c];  // a hanging DoF is never a parent (chain-free rows): no other warp reads v[h]
                    }
            }
        }
    }
}

// transpose: dst[col[k]] += w[k] dst[h]; dst[h] = 0; optionally src[h] = save[r]
__global__ void condense_kernel(uint32_t n_rows, const uint32_t *__restrict__ hdof, const uint32_t *__restrict__ ptr,
                                const uint32_t *__restrict__ col, const double *__restrict__ w, double *__restrict__ dst,
                                int ncomp, size_t stride, double *__restrict__ src, const double *__restrict__ save,
                                const int *__restrict__ skip)
{
    if (skip != nullptr && *skip != 0) return;
    const int lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rows; r += warps) {
        const uint32_t h = hdof[r], b = ptr[r], e = ptr[r + 1];
        for (int c0 = 0; c0 < ncomp; c0 += kCompChunk) {
            const int nc = min(kCompChunk, ncomp - c0);
            double t[kCompChunk];
#pragma unroll
            for (int c = 0; c < kCompChunk; ++c) t[c] = c < nc ? dst[(size_t)(c0 + c) * stride + h] : 0.0;
            for (uint32_t k = b + lane; k < e; k += 32) {
                const double wk = w[k];
                double *dk = dst + (size_t)c0 * stride + col[k];
#pragma unroll
                for (int c = 0; c < kCompChunk; ++c)
                    if (c < nc) atomicAdd(dk + c * stride, wk * t[c]);
            }
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int c = 0; c < kCompChunk; ++c)
                    if (c < nc) {
                        dst[(size_t)(c0 + c) * stride + h] = 0.0;
                        if (src) src[(size_t)(c0 + c) * stride + h] = save[(size_t)(c0 + c) * n_rows + r];
                    }
            }
        }
    }
}

// ---- face-structured constraints (b200fe_op_set_face_constraints) ---------------------------------------------------
// One CTA per coarse face: the (2p+1)^2 fine nodes of the face are W (x) W applied to its (p+1)^2 coarse nodes, so a face
// costs (p+1)^2 + (2p+1)^2 scattered accesses and two small 1-D contractions in shared memory instead of the
// (2p+1)^2 (p+1)^2 gathers of its CSR rows (the CSR kernels are L2-sector-bound).
// W[rel*nm + j] row-major [nf][nm]; parents [blk][a + nm*b]; children [blk][a' + nf*b'].
__global__ void distribute_faces_kernel(uint32_t n_blocks, int nm, int nf, const double *__restrict__ Wg, const uint32_t *__restrict__ parents,
                                        const uint32_t *__restrict__ children, double *__restrict__ v, int ncomp, size_t stride,
                                        double *__restrict__ save, const int *__restrict__ skip)
{
    if (skip != nullptr && *skip != 0) return;
    extern __shared__ double fsm[];
    double *W = fsm, *P = W + nf * nm, *T = P + nm * nm;  // T: [a' + nf*b]
    for (int t = threadIdx.x; t < nf * nm; t += blockDim.x) W[t] = Wg[t];
    for (uint32_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        const uint32_t *par = parents + (size_t)blk * nm * nm, *chi = children + (size_t)blk * nf * nf;
        for (int c = 0; c < ncomp; ++c) {
            double *vc = v + (size_t)c * stride;
            __syncthreads();  // W loaded / previous use of P and T finished
            for (int t = threadIdx.x; t < nm * nm; t += blockDim.x) {
                const uint32_t i = par[t];
                P[t] = i == kInvalidIndex ? 0.0 : vc[i];
            }
            __syncthreads();
            for (int t = threadIdx.x; t < nf * nm; t += blockDim.x) {
                const int a1 = t % nf, b = t / nf;
                double s = 0.0;
                for (int a = 0; a < nm; ++a) s = fma(W[a1 * nm + a], P[a + nm * b], s);
                T[t] = s;
            }
            __syncthreads();
            for (int t = threadIdx.x; t < nf * nf; t += blockDim.x) {
                const uint32_t i = chi[t];
                if (i == kInvalidIndex) continue;
                const int a1 = t % nf, b1 = t / nf;
                double s = 0.0;
                for (int b = 0; b < nm; ++b) s = fma(W[b1 * nm + b], T[a1 + nf * b], s);
                if (save) save[((size_t)c * n_blocks + blk) * nf * nf + t] = vc[i];
                vc[i] = s;  // every hanging DoF is the child of exactly one block and never a parent
            }
        }
    }
}

__global__ void condense_faces_kernel(uint32_t n_blocks, int nm, int nf, const double *__restrict__ Wg, const uint32_t *__restrict__ parents,
                                      const uint32_t *__restrict__ children, double *__restrict__ dst, int ncomp, size_t stride,
                                      double *__restrict__ src, const double *__restrict__ save, const int *__restrict__ skip)
{
    if (skip != nullptr && *skip != 0) return;
    extern __shared__ double fsm[];
    double *W = fsm, *C = W + nf * nm, *T = C + nf * nf;  // T: [a' + nf*b]
    for (int t = threadIdx.x; t < nf * nm; t += blockDim.x) W[t] = Wg[t];
    for (uint32_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        const uint32_t *par = parents + (size_t)blk * nm * nm, *chi = children + (size_t)blk * nf * nf;
        for (int c = 0; c < ncomp; ++c) {
            double *dc = dst + (size_t)c * stride;
            __syncthreads();
            for (int t = threadIdx.x; t < nf * nf; t += blockDim.x) {
                const uint32_t i = chi[t];
                double val = 0.0;
                if (i != kInvalidIndex) {
                    val = dc[i];
                    dc[i] = 0.0;
                    if (src) src[(size_t)c * stride + i] = save[((size_t)c * n_blocks + blk) * nf * nf + t];
                }
                C[t] = val;
            }
            __syncthreads();
            for (int t = threadIdx.x; t < nf * nm; t += blockDim.x) {
                const int a1 = t % nf, b = t / nf;
                double s = 0.0;
                for (int b1 = 0; b1 < nf; ++b1) s = fma(W[b1 * nm + b], C[a1 + nf * b1], s);
                T[t] = s;
            }
            __syncthreads();
            for (int t = threadIdx.x; t < nm * nm; t += blockDim.x) {
                const uint32_t i = par[t];
                if (i == kInvalidIndex) continue;
                const int a = t % nm, b = t / nm;
                double s = 0.0;
                for (int a1 = 0; a1 < nf; ++a1) s = fma(W[a1 * nm + a], T[a1 + nf * b], s);
                atomicAdd(dc + i, s);
            }
        }
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// operator pieces
// ---------------------------------------------------------------------------------------------
bool op_has_mc_kernel(const Operator &op)
{
    static const bool on = [] { const char *e = std::getenv("B200FE_MULTI_COMPONENT"); return !e || std::atoi(e) != 0; }();
    return on && op.collocated && op.qop == QOP_LAPLACE && op.otf_flag() == 0;
}

// Exclusive cell-interior DoFs: verified on the caller's index table, not assumed (a table that maps two cells' interior
// positions to one DoF keeps the atomic scatter everywhere).  Best effort: without the scratch memory the feature stays off.
static void op_prepare_exclusive(Operator &op)
{
    static const bool on = [] { const char *e = std::getenv("B200FE_EXCL_INTERIOR"); return !e || std::atoi(e) != 0; }();
    if (!on || op.nm < 4 || op.n_cells == 0 || op.n_local() == 0) return;  // p <= 2: thread-per-element kernels / no interior to speak of
    const uint32_t n_words = (op.n_local() + 31) / 32;
    uint32_t *d_bits = nullptr;  // seen | multi | violation word, then the mask that stays
    int h_violation = 1;
    if (cudaMalloc(&d_bits, (2 * (size_t)n_words + 1) * sizeof(uint32_t)) != cudaSuccess) { cudaGetLastError(); return; }
    if (cudaMalloc(&op.d_excl_mask, (size_t)n_words * sizeof(uint32_t)) != cudaSuccess) { cudaGetLastError(); cudaFree(d_bits); op.d_excl_mask = nullptr; return; }
    cudaMemset(d_bits, 0, (2 * (size_t)n_words + 1) * sizeof(uint32_t));
    cudaMemset(op.d_excl_mask, 0, (size_t)n_words * sizeof(uint32_t));
    const size_t n_entries = (size_t)op.n_cells * op.nm * op.nm * op.nm;
    excl_scan_kernel<<<148 * 16, 256>>>(n_entries, op.nm, op.d_idx, d_bits, d_bits + n_words, op.d_excl_mask);
    excl_check_kernel<<<148 * 4, 256>>>(n_words, d_bits + n_words, op.d_excl_mask, reinterpret_cast<int *>(d_bits + 2 * (size_t)n_words));
    g_launch_count += 2;
    const cudaError_t e = cudaMemcpy(&h_violation, d_bits + 2 * (size_t)n_words, sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(d_bits);
    if (e != cudaSuccess || h_violation != 0) {
        cudaGetLastError();
        cudaFree(op.d_excl_mask);
        op.d_excl_mask = nullptr;
    }
}

// the cell kernel of this operator: the sum-factorisation kernel family, or -- interpolated operator on axis-aligned cells
// -- the separable kernel on the nodal values (sumfact_cart.cuh)
static cudaError_t op_launch_kernel(Operator &op, const KArgs &a, cudaStream_t s, LaunchInfo *info, bool dry_run)
{
    if (op.cartesian && !op.collocated) return launch_cartesian(op.nm, op.qop, op.S.data(), a, s, info, dry_run);
    const int qop = op.qop | op.otf_flag();
    return launch_sumfact(op.nm, op.nq, op.collocated, qop, true, op.cartesian ? op.S.data() : op.B.data(), op.D.data(), a, s, info, dry_run,
                          op.otf_flag() ? op.W.data() : nullptr);
}

static CartVecs cart_vecs(const Operator &op)
{
    CartVecs v{};
    for (int i = 0; i < op.nm; ++i) {
        v.dk[i] = op.S[(size_t)i * op.nm + i];
        v.dm[i] = op.S[(size_t)op.nm * op.nm + (size_t)i * op.nm + i];
        v.mv[i] = op.mvec[i];
    }
    return v;
}

int op_apply_cells(Operator &op, double *d_dst, const double *d_src, uint32_t cb, uint32_t ce,
                   double *d_dot, cudaStream_t s, int ncomp)
{
    if (ce <= cb) return B200FE_OK;
    const size_t nm3 = (size_t)op.nm * op.nm * op.nm, nq3 = (size_t)op.nq * op.nq * op.nq;
    const bool affine = op.otf_flag() != 0;  // geometric factors evaluated on the fly (affine or trilinear cells)
    KArgs a{ce - cb, (op.d_G && !affine) ? op.d_G + cb * 6 * nq3 : nullptr, op.d_JxW ? op.d_JxW + cb * nq3 : nullptr,
            d_src, d_dst, op.d_idx + cb * nm3, d_dot, affine ? op.otf_data() + (size_t)cb * op.otf_stride() : nullptr, op.d_skip};
    a.ncomp = ncomp;
    a.comp_stride = op.n_local();
    a.excl_interior = op.d_excl_mask != nullptr;
    const bool timed = op.timing && op.ev_used + 2 <= op.ev.size();
    if (timed) B200FE_CUDA_TRY(cudaEventRecord(op.ev[op.ev_used], s));
    B200FE_CUDA_TRY(op_launch_kernel(op, a, s, &op.last_launch, false));
    if (timed) {
        B200FE_CUDA_TRY(cudaEventRecord(op.ev[op.ev_used + 1], s));
        op.ev_used += 2;
    }
    ++g_launch_count;
    return B200FE_OK;
}

int op_copy_constrained(Operator &op, double *d_dst, const double *d_src, double *d_dot, cudaStream_t s, int ncomp)
{
    if (op.n_constrained == 0) return B200FE_OK;
    const dim3 blocks(std::min<unsigned>((op.n_constrained + 255) / 256, 1184), (unsigned)ncomp);
    copy_constrained_kernel<<<blocks, 256, 0, s>>>(op.n_constrained, op.d_constrained, op.n_local(), d_src, d_dst, d_dot, op.d_skip);
    B200FE_CUDA_TRY(cudaGetLastError());
    ++g_launch_count;
    return B200FE_OK;
}

// the save buffer holds n_hang doubles per component; grown on first use with more components
static int ensure_hang_save(Operator &op, int ncomp)
{
    if (ncomp <= op.hang_save_comps) return B200FE_OK;
    cudaFree(op.d_hang_save);
    op.d_hang_save = nullptr;
    op.hang_save_comps = 0;
    B200FE_CUDA_TRY(cudaMalloc(&op.d_hang_save, (size_t)op.n_hang * ncomp * sizeof(double)));
    op.hang_save_comps = ncomp;
    return B200FE_OK;
}

static int ensure_face_save(Operator &op, int ncomp)
{
    if (ncomp <= op.face_save_comps) return B200FE_OK;
    cudaFree(op.d_face_save);
    op.d_face_save = nullptr;
    op.face_save_comps = 0;
    B200FE_CUDA_TRY(cudaMalloc(&op.d_face_save, (size_t)op.n_face_blocks * op.face_nf * op.face_nf * ncomp * sizeof(double)));
    op.face_save_comps = ncomp;
    return B200FE_OK;
}

static size_t face_smem(const Operator &op)
{
    return sizeof(double) * (size_t)(2 * op.face_nf * op.face_nm + op.face_nf * op.face_nf);  // W + max(P, C) + T
}

int op_distribute(Operator &op, double *d_v, bool save, cudaStream_t s, int ncomp)
{
    if (op.n_face_blocks) {  // face-structured form takes precedence over the CSR rows
        if (save)
            if (int rc = ensure_face_save(op, ncomp)) return rc;
        distribute_faces_kernel<<<std::min<unsigned>(op.n_face_blocks, 148u * 8u), 128, face_smem(op), s>>>(
            op.n_face_blocks, op.face_nm, op.face_nf, op.d_face_W, op.d_face_parents, op.d_face_children, d_v, ncomp, op.n_local(),
            save ? op.d_face_save : nullptr, op.d_skip);
        B200FE_CUDA_TRY(cudaGetLastError());
        ++g_launch_count;
        return B200FE_OK;
    }
    if (op.n_hang == 0) return B200FE_OK;
    if (save)
        if (int rc = ensure_hang_save(op, ncomp)) return rc;
    const unsigned blocks = std::min<unsigned>((op.n_hang + 7) / 8, 148u * 8u);  // 8 warps (rows) per CTA
    distribute_kernel<<<blocks, 256, 0, s>>>(op.n_hang, op.d_hang_dof, op.d_hang_ptr, op.d_hang_col, op.d_hang_w, d_v, ncomp,
                                             op.n_local(), save ? op.d_hang_save : nullptr, op.d_skip);
    B200FE_CUDA_TRY(cudaGetLastError());
    ++g_launch_count;
    return B200FE_OK;
}

int op_condense(Operator &op, double *d_dst, double *d_src_restore, cudaStream_t s, int ncomp)
{
    if (op.n_face_blocks) {
        condense_faces_kernel<<<std::min<unsigned>(op.n_face_blocks, 148u * 8u), 128, face_smem(op), s>>>(
            op.n_face_blocks, op.face_nm, op.face_nf, op.d_face_W, op.d_face_parents, op.d_face_children, d_dst, ncomp, op.n_local(),
            d_src_restore, op.d_face_save, op.d_skip);
        B200FE_CUDA_TRY(cudaGetLastError());
        ++g_launch_count;
        return B200FE_OK;
    }
    if (op.n_hang == 0) return B200FE_OK;
    const unsigned blocks = std::min<unsigned>((op.n_hang + 7) / 8, 148u * 8u);
    condense_kernel<<<blocks, 256, 0, s>>>(op.n_hang, op.d_hang_dof, op.d_hang_ptr, op.d_hang_col, op.d_hang_w, d_dst, ncomp,
                                           op.n_local(), d_src_restore, op.d_hang_save, op.d_skip);
    B200FE_CUDA_TRY(cudaGetLastError());
    ++g_launch_count;
    return B200FE_OK;
}

int op_vmult(Operator &op, double *d_dst, const double *d_src, double *d_dot, bool ghost_on, bool compute_on,
             cudaStream_t s, int ncomp, bool dst_is_zero)
{
    NvtxRange range("matvec");
    Halo *h = op.halo;
    const size_t stride = op.n_local();
    double *src_mut = const_cast<double *>(d_src);  // ghost (and hanging) entries of src are scratch, as in deal.II
    // hanging-node rows need the parents' ghost values before the first cell runs: no overlap split with constraints
    const bool split = h && ghost_on && compute_on && (op.n_phase0 + op.n_phase1 > 0) && !op.has_constraints();
    if (split && ncomp > 1) {  // the overlap schedule is per vector: one component after the other
        for (int c = 0; c < ncomp; ++c)
            if (int rc = op_vmult(op, d_dst + c * stride, d_src + c * stride, d_dot, ghost_on, compute_on, s, 1, dst_is_zero)) return rc;
        return B200FE_OK;
    }
    if (compute_on && !dst_is_zero) B200FE_CUDA_TRY(cudaMemsetAsync(d_dst, 0, sizeof(double) * stride * ncomp, s));
    if (split) {
        // 3-phase overlap (bakeoff_problems_dealii/include/portable_laplace_operator.h:643-696)
        if (int rc = halo_update_ghosts_start(*h, src_mut, s)) return rc;
        if (int rc = op_apply_cells(op, d_dst, d_src, 0, op.n_phase0, d_dot, s)) return rc;
        if (int rc = halo_update_ghosts_finish(*h, s)) return rc;
        if (int rc = op_apply_cells(op, d_dst, d_src, op.n_phase0, op.n_phase0 + op.n_phase1, d_dot, s)) return rc;
        if (int rc = halo_compress_start(*h, d_dst, s)) return rc;
        if (int rc = op_apply_cells(op, d_dst, d_src, op.n_phase0 + op.n_phase1, op.n_cells, d_dot, s)) return rc;
        if (int rc = halo_compress_finish(*h, d_dst, s)) return rc;
        if (int rc = halo_zero_ghosts(*h, src_mut, s)) return rc;
        return op_copy_constrained(op, d_dst, d_src, d_dot, s, 1);
    }
    // vector-valued: the scalar operator on every component; constraint rows and identity rows are applied to all
    // components in one launch each (their index lists are read once)
    if (h && ghost_on)
        if (int rc = halo_update_ghosts_components(*h, src_mut, ncomp, stride, s)) return rc;
    if (compute_on) {
        if (int rc = op_distribute(op, src_mut, true, s, ncomp)) return rc;
        if (ncomp > 1 && op_has_mc_kernel(op)) {  // one launch: the geometric factors of a cell serve all components
            if (int rc = op_apply_cells(op, d_dst, d_src, 0, op.n_cells, d_dot, s, ncomp)) return rc;
        } else
            for (int c = 0; c < ncomp; ++c)
                if (int rc = op_apply_cells(op, d_dst + c * stride, d_src + c * stride, 0, op.n_cells, d_dot, s)) return rc;
        if (int rc = op_condense(op, d_dst, src_mut, s, ncomp)) return rc;
    }
    if (h && ghost_on) {
        if (int rc = halo_compress_add_components(*h, d_dst, ncomp, stride, s)) return rc;
        for (int c = 0; c < ncomp; ++c)
            if (int rc = halo_zero_ghosts(*h, src_mut + c * stride, s)) return rc;
    }
    if (ghost_on)  // reference: copy_constrained_values sits inside the ghost_exchange_on branch (:229-234)
        if (int rc = op_copy_constrained(op, d_dst, d_src, d_dot, s, ncomp)) return rc;
    return B200FE_OK;
}

namespace {
struct SetupScratch {  // setup-time device scratch, freed on every exit path
    std::vector<void *> p;
    ~SetupScratch() { for (void *q : p) cudaFree(q); }
    template <class T> cudaError_t put(T **d, const std::vector<T> &h)
    {
        cudaError_t e = cudaMalloc(d, std::max<size_t>(h.size(), 1) * sizeof(T));
        if (e != cudaSuccess) return e;
        p.push_back(*d);
        return h.empty() ? cudaSuccess : cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    }
    template <class T> cudaError_t make(T **d, size_t n)
    {
        cudaError_t e = cudaMalloc(d, std::max<size_t>(n, 1) * sizeof(T));
        if (e == cudaSuccess) p.push_back(*d);
        return e;
    }
};
}  // namespace

// diag(C^T A C): plain cell diagonals on the cells without hanging DoFs, local C_c^T A_c C_c on the others (setup path).
int op_diagonal_constrained(Operator &op, double *d_diag, cudaStream_t s)
{
    const int nm3 = op.nm * op.nm * op.nm, nq3 = op.nq * op.nq * op.nq;
    const uint32_t n_local = op.n_local();
    // host view of the index table and of the rows
    std::vector<uint32_t> idx((size_t)op.n_cells * nm3);
    B200FE_CUDA_TRY(cudaMemcpyAsync(idx.data(), op.d_idx, idx.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    B200FE_CUDA_TRY(cudaStreamSynchronize(s));
    std::vector<int32_t> row_of(n_local, -1);
    for (uint32_t r = 0; r < op.n_hang; ++r) row_of[op.h_hang_dof[r]] = (int32_t)r;
    std::vector<uint8_t> is_if(op.n_cells, 0);
    std::vector<uint32_t> cells_if;
    for (uint32_t c = 0; c < op.n_cells; ++c) {
        bool any = false;
        for (int a = 0; a < nm3 && !any; ++a) {
            const uint32_t id = idx[(size_t)c * nm3 + a];
            any = id != kInvalidIndex && row_of[id] >= 0;
        }
        if (any) { is_if[c] = 1; cells_if.push_back(c); }
    }
    const uint32_t n_if = (uint32_t)cells_if.size();
    // local constraint matrices C_c: per (cell, local a) a run of (compact target id, weight); targets per cell deduplicated
    std::vector<uint32_t> ent_ptr((size_t)n_if * nm3 + 1, 0), tgt_ptr(n_if + 1, 0), tgt, idx_if((size_t)n_if * nm3);
    std::vector<uint16_t> ent_cid;
    std::vector<double> ent_w;
    std::vector<int32_t> cid_of(n_local, -1);
    uint32_t max_targets = 1;
    for (uint32_t i = 0; i < n_if; ++i) {
        const uint32_t c = cells_if[i];
        const uint32_t t0 = (uint32_t)tgt.size();
        auto cid = [&](uint32_t j) -> uint16_t {
            if (cid_of[j] < 0) { cid_of[j] = (int32_t)(tgt.size() - t0); tgt.push_back(j); }
            return (uint16_t)cid_of[j];
        };
        for (int a = 0; a < nm3; ++a) {
            const uint32_t id = idx[(size_t)c * nm3 + a];
            idx_if[(size_t)i * nm3 + a] = id == kInvalidIndex ? kInvalidIndex : (uint32_t)((size_t)i * nm3 + a);
            if (id != kInvalidIndex) {
                if (row_of[id] < 0) { ent_cid.push_back(cid(id)); ent_w.push_back(1.0); }
                else
                    for (uint32_t k = op.h_hang_ptr[row_of[id]]; k < op.h_hang_ptr[row_of[id] + 1]; ++k) {
                        ent_cid.push_back(cid(op.h_hang_col[k])); ent_w.push_back(op.h_hang_w[k]);
                    }
            }
            ent_ptr[(size_t)i * nm3 + a + 1] = (uint32_t)ent_cid.size();
        }
        for (uint32_t k = t0; k < tgt.size(); ++k) cid_of[tgt[k]] = -1;
        tgt_ptr[i + 1] = (uint32_t)tgt.size();
        max_targets = std::max<uint32_t>(max_targets, (uint32_t)tgt.size() - t0);
        if (tgt.size() - t0 > 0xFFFFu) return fail(B200FE_ERR_UNSUPPORTED, "b200fe_op_diagonal: more than 65535 constraint targets in one cell");
    }
    if ((size_t)n_if * nm3 >= 0xFFFFFFFFull) return fail(B200FE_ERR_UNSUPPORTED, "b200fe_op_diagonal: too many interface cells");
    if (max_targets * sizeof(double) > 200 * 1024) return fail(B200FE_ERR_UNSUPPORTED, "b200fe_op_diagonal: constraint targets of one cell exceed shared memory");

    SetupScratch dev;
    uint8_t *d_is_if = nullptr;
    B200FE_CUDA_TRY(dev.put(&d_is_if, is_if));
    B200FE_CUDA_TRY(cudaMemsetAsync(d_diag, 0, sizeof(double) * n_local, s));
    if (op.n_cells) {
        diagonal_kernel<<<std::min<uint32_t>(op.n_cells, 148u * 8u), 128, 0, s>>>(op.n_cells, op.nm, op.nq, op.qop, op.d_mats, op.d_G, op.d_JxW, op.d_idx, d_diag, d_is_if);
        B200FE_CUDA_TRY(cudaGetLastError());
    }
    if (n_if) {
        uint32_t *d_cells = nullptr, *d_ent_ptr = nullptr, *d_tgt_ptr = nullptr, *d_tgt = nullptr, *d_idx_if = nullptr;
        uint16_t *d_ent_cid = nullptr;
        double *d_ent_w = nullptr, *d_G = nullptr, *d_J = nullptr, *d_src = nullptr, *d_dst = nullptr;
        B200FE_CUDA_TRY(dev.put(&d_cells, cells_if));
        B200FE_CUDA_TRY(dev.put(&d_ent_ptr, ent_ptr));
        B200FE_CUDA_TRY(dev.put(&d_ent_cid, ent_cid));
        B200FE_CUDA_TRY(dev.put(&d_ent_w, ent_w));
        B200FE_CUDA_TRY(dev.put(&d_tgt_ptr, tgt_ptr));
        B200FE_CUDA_TRY(dev.put(&d_tgt, tgt));
        B200FE_CUDA_TRY(dev.put(&d_idx_if, idx_if));
        B200FE_CUDA_TRY(dev.make(&d_src, (size_t)n_if * nm3));
        B200FE_CUDA_TRY(dev.make(&d_dst, (size_t)n_if * nm3));
        const unsigned gb = 148u * 8u;
        if (op.qop & QOP_LAPLACE) {
            B200FE_CUDA_TRY(dev.make(&d_G, (size_t)n_if * 6 * nq3));
            gather_cells_kernel<<<gb, 256, 0, s>>>(n_if, d_cells, (size_t)6 * nq3, op.d_G, d_G);
        }
        if (op.qop & QOP_MASS) {
            B200FE_CUDA_TRY(dev.make(&d_J, (size_t)n_if * nq3));
            gather_cells_kernel<<<gb, 256, 0, s>>>(n_if, d_cells, (size_t)nq3, op.d_JxW, d_J);
        }
        B200FE_CUDA_TRY(cudaGetLastError());
        const size_t red_smem = max_targets * sizeof(double);
        if (red_smem > 48 * 1024)
            B200FE_CUDA_TRY(cudaFuncSetAttribute(diag_constrained_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)red_smem));
        KArgs a{n_if, d_G, d_J, d_src, d_dst, d_idx_if, nullptr, nullptr, nullptr};
        for (int b = 0; b < nm3; ++b) {
            unit_vectors_kernel<<<std::min<unsigned>((unsigned)(((size_t)n_if * nm3 + 255) / 256), gb), 256, 0, s>>>(n_if, nm3, b, d_idx_if, d_src, d_dst);
            B200FE_CUDA_TRY(launch_sumfact(op.nm, op.nq, op.collocated, op.qop, true, op.B.data(), op.D.data(), a, s, nullptr, false));
            diag_constrained_reduce_kernel<<<std::min<unsigned>(n_if, gb), 128, red_smem, s>>>(n_if, nm3, b, d_dst, d_ent_ptr, d_ent_cid, d_ent_w, d_tgt_ptr,
                                                                                              d_tgt, d_diag);
            B200FE_CUDA_TRY(cudaGetLastError());
        }
    }
    if (op.halo)
        if (int rc = halo_compress_add(*op.halo, d_diag, s)) return rc;
    // constrained rows (Dirichlet and hanging) act as identity in the preconditioner: 1 on the diagonal
    if (op.n_constrained) {
        set_constrained_kernel<<<std::min<unsigned>((op.n_constrained + 255) / 256, 1184), 256, 0, s>>>(op.n_constrained, op.d_constrained, 1.0, d_diag);
        B200FE_CUDA_TRY(cudaGetLastError());
    }
    set_constrained_kernel<<<std::min<unsigned>((op.n_hang + 255) / 256, 1184), 256, 0, s>>>(op.n_hang, op.d_hang_dof, 1.0, d_diag);
    B200FE_CUDA_TRY(cudaGetLastError());
    B200FE_CUDA_TRY(cudaStreamSynchronize(s));  // the scratch arrays are freed on return
    return B200FE_OK;
}

}  // namespace b200fe

using namespace b200fe;

extern "C" {

// shared body of b200fe_boxmesh_nodes / b200fe_hangmesh_nodes: cell table [n_cells][stride] -> mapping support points
static int nodes_from_cell_table(const std::vector<int32_t> &cells, int stride, uint32_t n_cells, const double p1[3],
                                 const double hh[3], int p_geo, int deform_kind, double amplitude, double frequency,
                                 double *d_nodes, cudaStream_t s)
{
    const int ng = p_geo + 1;
    std::vector<double> t(ng), w(ng);
    if (int rc = b200fe_basis_1d(p_geo, ng, B200FE_QUAD_GLL, nullptr, nullptr, nullptr, t.data(), w.data())) return rc;
    int32_t *d_xyz = nullptr;
    double *d_t = nullptr;
    B200FE_CUDA_TRY(cudaMalloc(&d_xyz, cells.size() * sizeof(int32_t)));
    cudaError_t e = cudaMalloc(&d_t, ng * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_xyz, cells.data(), cells.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_t, t.data(), ng * sizeof(double), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
        const uint64_t total = (uint64_t)n_cells * ng * ng * ng;
        box_nodes_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(n_cells, ng, d_xyz, stride, p1[0], p1[1], p1[2], hh[0], hh[1],
                                                                         hh[2], d_t, deform_kind, amplitude, frequency, d_nodes);
        e = cudaGetLastError();
    }
    cudaStreamSynchronize(s);  // setup call: the temporaries are freed right away
    cudaFree(d_xyz);
    cudaFree(d_t);
    if (e != cudaSuccess) return fail_cuda(e, "box_nodes_kernel");
    return B200FE_OK;
}

int b200fe_boxmesh_nodes(const b200fe_boxmesh *mesh, int p_geo, int deform_kind, double amplitude,
                         double frequency, double *d_nodes, void *stream)
{
    B200FE_REQUIRE(mesh && d_nodes, "b200fe_boxmesh_nodes: null pointer");
    B200FE_REQUIRE(p_geo >= 1 && p_geo <= 8, "b200fe_boxmesh_nodes: p_geo outside 1..8");
    B200FE_REQUIRE(deform_kind == 0 || deform_kind == 1, "b200fe_boxmesh_nodes: unknown deformation");
    b200fe_boxmesh_info_t info;
    if (int rc = b200fe_boxmesh_info(mesh, &info)) return rc;
    if (info.n_cells_local == 0) return B200FE_OK;
    std::vector<int32_t> xyz((size_t)info.n_cells_local * 3);
    if (int rc = b200fe_boxmesh_fill(mesh, nullptr, nullptr, nullptr, nullptr, xyz.data(), nullptr)) return rc;
    return nodes_from_cell_table(xyz, 3, info.n_cells_local, info.origin, info.h, p_geo, deform_kind, amplitude, frequency, d_nodes,
                                 (cudaStream_t)stream);
}

int b200fe_boxmesh_dof_indices_device(const b200fe_boxmesh *mesh, uint32_t *d_dof_indices, void *stream)
{
    B200FE_REQUIRE(mesh && d_dof_indices, "b200fe_boxmesh_dof_indices_device: null pointer");
    BoxMeshTables t;
    if (int rc = boxmesh_tables(mesh, &t)) return rc;
    if (t.n_cells_local == 0) return B200FE_OK;
    const int nm = t.p + 1, nm3 = nm * nm * nm;
    std::vector<uint32_t> ent_idx(nm3);
    for (int l = 0; l < nm3; ++l) ent_idx[l] = (uint32_t)t.l_ent[l] << 16 | (uint32_t)t.l_idx[l];
    cudaStream_t s = (cudaStream_t)stream;
    SetupScratch scratch;
    uint32_t *d_lbase = nullptr, *d_ei = nullptr;
    int32_t *d_xyz = nullptr;
    B200FE_CUDA_TRY(scratch.make(&d_lbase, (size_t)t.n_cells_local * 27));
    B200FE_CUDA_TRY(scratch.make(&d_xyz, (size_t)t.n_cells_local * 3));
    B200FE_CUDA_TRY(scratch.make(&d_ei, (size_t)nm3));
    B200FE_CUDA_TRY(cudaMemcpyAsync(d_lbase, t.lbase, (size_t)t.n_cells_local * 27 * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    B200FE_CUDA_TRY(cudaMemcpyAsync(d_xyz, t.cell_xyz, (size_t)t.n_cells_local * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    B200FE_CUDA_TRY(cudaMemcpyAsync(d_ei, ent_idx.data(), (size_t)nm3 * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    const size_t n_entries = (size_t)t.n_cells_local * nm3;
    expand_indices_kernel<<<(unsigned)std::min<size_t>((n_entries + 255) / 256, 148u * 32u), 256, 0, s>>>(
        n_entries, nm, t.p, t.dirichlet, t.cells[0] * t.p, t.cells[1] * t.p, t.cells[2] * t.p, d_lbase, d_xyz, d_ei, d_dof_indices);
    B200FE_CUDA_TRY(cudaGetLastError());
    ++g_launch_count;
    B200FE_CUDA_TRY(cudaStreamSynchronize(s));  // setup call: the temporaries are freed on return
    return B200FE_OK;
}

int b200fe_hangmesh_nodes(const b200fe_hangmesh *mesh, int p_geo, int deform_kind, double amplitude,
                          double frequency, double *d_nodes, void *stream)
{
    B200FE_REQUIRE(mesh && d_nodes, "b200fe_hangmesh_nodes: null pointer");
    B200FE_REQUIRE(p_geo >= 1 && p_geo <= 8, "b200fe_hangmesh_nodes: p_geo outside 1..8");
    B200FE_REQUIRE(deform_kind == 0 || deform_kind == 1, "b200fe_hangmesh_nodes: unknown deformation");
    b200fe_hangmesh_info_t info;
    if (int rc = b200fe_hangmesh_info(mesh, &info)) return rc;
    if (info.n_cells_local == 0) return B200FE_OK;
    std::vector<int32_t> lxyz((size_t)info.n_cells_local * 4);
    if (int rc = b200fe_hangmesh_fill(mesh, nullptr, nullptr, nullptr, nullptr, lxyz.data(), nullptr, nullptr, nullptr, nullptr, nullptr)) return rc;
    return nodes_from_cell_table(lxyz, 4, info.n_cells_local, info.origin, info.h, p_geo, deform_kind, amplitude, frequency, d_nodes,
                                 (cudaStream_t)stream);
}

int b200fe_geometry_from_nodes(int p_geo, int nq, int quad_kind, uint32_t n_cells, const double *d_nodes,
                               double *d_G, double *d_JxW, void *stream)
{
    B200FE_REQUIRE(p_geo >= 1 && p_geo <= 8, "b200fe_geometry_from_nodes: p_geo outside 1..8");
    B200FE_REQUIRE(nq >= 2 && nq <= 10, "b200fe_geometry_from_nodes: nq outside 2..10");
    B200FE_REQUIRE(n_cells == 0 || (d_nodes && (d_G || d_JxW)), "b200fe_geometry_from_nodes: null pointer");
    if (n_cells == 0) return B200FE_OK;
    const int ng = p_geo + 1;
    // geometry basis (degree p_geo on GLL nodes) at the operator's quadrature points
    std::vector<double> sv(ng * nq), sg(ng * nq), w(nq);
    if (int rc = b200fe_basis_1d(p_geo, nq, quad_kind, nullptr, nullptr, sg.data(), nullptr, w.data())) return rc;
    {   // values: always the true interpolation matrix (not the collocation identity shortcut)
        std::vector<double> xq(nq), nodes(ng), wn(ng), V(nq * ng);
        b200fe_basis_1d(p_geo, nq, quad_kind, nullptr, nullptr, nullptr, xq.data(), nullptr);
        b200fe_basis_1d(p_geo, ng, B200FE_QUAD_GLL, nullptr, nullptr, nullptr, nodes.data(), wn.data());
        for (int q = 0; q < nq; ++q)
            for (int a = 0; a < ng; ++a) {
                double v = 1.0;
                for (int m = 0; m < ng; ++m)
                    if (m != a) v *= (xq[q] - nodes[m]) / (nodes[a] - nodes[m]);
                sv[a * nq + q] = v;
            }
    }
    GeoMats gm;
    std::memset(&gm, 0, sizeof(gm));
    for (int q = 0; q < nq; ++q) {
        gm.w[q] = w[q];
        for (int a = 0; a < ng; ++a) {
            gm.V[q * ng + a] = sv[a * nq + q];
            gm.dV[q * ng + a] = sg[a * nq + q];
        }
    }
    const int threads = std::min(256, ((nq * nq * nq + 31) / 32) * 32);
    const unsigned blocks = std::min<uint32_t>(n_cells, 148u * 16u);
    geometry_kernel<<<blocks, threads, 3 * ng * ng * ng * sizeof(double), (cudaStream_t)stream>>>(gm, n_cells, ng, nq, d_nodes, d_G, d_JxW);
    B200FE_CUDA_TRY(cudaGetLastError());
    return B200FE_OK;
}

int b200fe_geometry_from_inv_jacobian(uint32_t n_cells, int nq, const double *d_inv_jacobian, const double *d_JxW, double *d_G,
                                      void *stream)
{
    B200FE_REQUIRE(nq >= 2 && nq <= 10, "b200fe_geometry_from_inv_jacobian: nq outside 2..10");
    B200FE_REQUIRE(n_cells == 0 || (d_inv_jacobian && d_JxW && d_G), "b200fe_geometry_from_inv_jacobian: null pointer");
    if (n_cells == 0) return B200FE_OK;
    const size_t total = (size_t)n_cells * nq * nq * nq;
    const unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, 148u * 16u);
    g_from_inv_jacobian_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(n_cells, nq * nq * nq, d_inv_jacobian, d_JxW, d_G);
    B200FE_CUDA_TRY(cudaGetLastError());
    return B200FE_OK;
}

int b200fe_geometry_affine_from_nodes(uint32_t n_cells, const double *d_nodes, double *d_cell_G, void *stream)
{
    B200FE_REQUIRE(n_cells == 0 || (d_nodes && d_cell_G), "b200fe_geometry_affine_from_nodes: null pointer");
    if (n_cells == 0) return B200FE_OK;
    affine_geometry_kernel<<<(n_cells + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n_cells, d_nodes, d_cell_G);
    B200FE_CUDA_TRY(cudaGetLastError());
    return B200FE_OK;
}

int b200fe_op_create(const b200fe_op_desc *d, b200fe_op **out)
{
    B200FE_REQUIRE(d && out, "b200fe_op_create: null pointer");
    if (d->p < 1 || d->p > 8) return fail(B200FE_ERR_UNSUPPORTED, "b200fe_op_create: degree p=%d outside 1..8", d->p);
    const int nm = d->p + 1;
    B200FE_REQUIRE(d->op_kind >= 1 && d->op_kind <= 3, "b200fe_op_create: op_kind must be LAPLACE, MASS or HELMHOLTZ");
    const bool coll = d->collocated != 0;
    // built variants (inst.cu): Laplace nq=p+2 | p+1 | collocated, mass nq=p+2, Helmholtz nq=p+1
    bool ok = false;
    if (d->op_kind == B200FE_OP_LAPLACE) ok = coll ? d->nq == nm : (d->nq == nm || d->nq == nm + 1);
    if (d->op_kind == B200FE_OP_MASS) ok = !coll && d->nq == nm + 1;
    if (d->op_kind == B200FE_OP_HELMHOLTZ) ok = !coll && d->nq == nm;
    if (!ok) return fail(B200FE_ERR_UNSUPPORTED, "b200fe_op_create: no kernel for op_kind=%d p=%d nq=%d collocated=%d", d->op_kind, d->p, d->nq, (int)coll);
    B200FE_REQUIRE(d->h_co_shape_gradients && (coll || d->h_shape_values), "b200fe_op_create: 1-D matrices missing");
    B200FE_REQUIRE(d->n_cells == 0 || d->d_dof_indices, "b200fe_op_create: dof_indices missing");
    const bool trilinear = d->d_cell_vertices != nullptr;
    const bool affine = d->d_cell_G != nullptr || trilinear;  // on-the-fly geometry of either kind
    B200FE_REQUIRE(!(d->d_cell_G && trilinear), "b200fe_op_create: d_cell_G and d_cell_vertices are mutually exclusive");
    if (trilinear && d->op_kind != B200FE_OP_LAPLACE)
        return fail(B200FE_ERR_UNSUPPORTED, "b200fe_op_create: trilinear on-the-fly geometry is built for the Laplace operators only");
    // (mass / Helmholtz with d_cell_G: axis-aligned cells only -- checked below, once the constants have been looked at)
    B200FE_REQUIRE(!affine || d->h_weights, "b200fe_op_create: on-the-fly geometry needs the 1-D quadrature weights");
    B200FE_REQUIRE(!trilinear || d->h_points, "b200fe_op_create: trilinear geometry needs the 1-D quadrature points");
    B200FE_REQUIRE(!(d->op_kind & B200FE_OP_LAPLACE) || d->n_cells == 0 || d->d_G || affine, "b200fe_op_create: G missing");
    B200FE_REQUIRE(!(d->op_kind & B200FE_OP_MASS) || d->n_cells == 0 || d->d_JxW || d->d_cell_G, "b200fe_op_create: JxW missing");
    B200FE_REQUIRE(d->n_constrained == 0 || d->h_constrained, "b200fe_op_create: constrained list missing");
    B200FE_REQUIRE((uint64_t)d->n_phase0 + d->n_phase1 <= d->n_cells, "b200fe_op_create: phase split exceeds n_cells");

    auto op = std::make_unique<Operator>();
    op->p = d->p; op->nm = nm; op->nq = d->nq; op->qop = d->op_kind; op->collocated = coll;
    op->n_cells = d->n_cells; op->n_owned = d->n_owned; op->n_ghost = d->n_ghost; op->n_constrained = d->n_constrained;
    op->n_phase0 = d->n_phase0; op->n_phase1 = d->n_phase1;
    op->d_idx = d->d_dof_indices; op->d_G = d->d_G; op->d_JxW = d->d_JxW; op->d_cellG = d->d_cell_G; op->d_cellX = d->d_cell_vertices;
    if (affine) op->W.assign(d->h_weights, d->h_weights + d->nq);
    if (trilinear) op->W.insert(op->W.end(), d->h_points, d->h_points + d->nq);
    const int nq = d->nq;
    op->shape_values.assign(nm * nq, 0.0);
    if (coll) for (int i = 0; i < nm; ++i) op->shape_values[i * nq + i] = 1.0;
    else op->shape_values.assign(d->h_shape_values, d->h_shape_values + nm * nq);
    op->co_shape_gradients.assign(d->h_co_shape_gradients, d->h_co_shape_gradients + nq * nq);
    // kernel (BK) layouts: B[q*nm+i] = shape_values[i*nq+q]; D[p*nq+n] = co_shape_gradients[n*nq+p]
    op->B.resize(nq * nm); op->D.resize(nq * nq);
    for (int q = 0; q < nq; ++q) {
        for (int i = 0; i < nm; ++i) op->B[q * nm + i] = op->shape_values[i * nq + q];
        for (int n = 0; n < nq; ++n) op->D[q * nq + n] = op->co_shape_gradients[n * nq + q];
    }
    // setup-kernel matrices on the device; shape gradients = D * B exactly (degree p <= nq-1)
    std::vector<double> mats(nm * nq + nq * nq + nm * nq);
    std::memcpy(mats.data(), op->shape_values.data(), sizeof(double) * nm * nq);
    std::memcpy(mats.data() + nm * nq, op->co_shape_gradients.data(), sizeof(double) * nq * nq);
    for (int i = 0; i < nm; ++i)
        for (int q = 0; q < nq; ++q) {
            double s = 0.0;
            for (int n = 0; n < nq; ++n) s += op->D[q * nq + n] * op->shape_values[i * nq + n];
            mats[nm * nq + nq * nq + i * nq + q] = s;
        }
    B200FE_CUDA_TRY(cudaMalloc(&op->d_mats, mats.size() * sizeof(double)));
    B200FE_CUDA_TRY(cudaMemcpy(op->d_mats, mats.data(), mats.size() * sizeof(double), cudaMemcpyHostToDevice));
    if (d->n_constrained) {
        B200FE_CUDA_TRY(cudaMalloc(&op->d_constrained, d->n_constrained * sizeof(uint32_t)));
        B200FE_CUDA_TRY(cudaMemcpy(op->d_constrained, d->h_constrained, d->n_constrained * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    if (d->d_cell_G && d->n_cells) {
        // axis-aligned cells (deal.II's "cartesian" cell type): the separable kernel.  B200FE_CARTESIAN=0 keeps the general
        // affine kernel (read at every create: tests switch it).
        const char *e = std::getenv("B200FE_CARTESIAN");
        if (!e || std::atoi(e) != 0) {
            int *d_flag = nullptr, h_flag = 1;
            B200FE_CUDA_TRY(cudaMalloc(&d_flag, sizeof(int)));
            cudaMemset(d_flag, 0, sizeof(int));
            cartesian_check_kernel<<<148 * 4, 256>>>(d->n_cells, d->d_cell_G, d_flag);
            ++g_launch_count;
            const cudaError_t ce = cudaMemcpy(&h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost);
            cudaFree(d_flag);
            if (ce != cudaSuccess) return fail_cuda(ce, "cartesian_check_kernel");
            if (h_flag == 0) {
                // 1-D stiffness and mass matrices on the nodal basis, K = (D B)^T W (D B), M = B^T W B (collocated: B = I, so
                // K = D^T W D and M = diag(w)), and m = B^T w; S = K | M, nm*nm doubles each
                op->cartesian = true;
                std::vector<double> DB((size_t)nq * nm, 0.0);  // derivative of shape i at point p
                for (int pp = 0; pp < nq; ++pp)
                    for (int i = 0; i < nm; ++i) {
                        double t = 0.0;
                        for (int n = 0; n < nq; ++n) t += op->D[pp * nq + n] * op->B[n * nm + i];
                        DB[(size_t)pp * nm + i] = t;
                    }
                op->S.assign(2 * (size_t)nm * nm, 0.0);
                op->mvec.assign(nm, 0.0);
                for (int i = 0; i < nm; ++i) {
                    for (int j = 0; j < nm; ++j) {
                        double k = 0.0, mm = 0.0;
                        for (int pp = 0; pp < nq; ++pp) {
                            k += DB[(size_t)pp * nm + i] * d->h_weights[pp] * DB[(size_t)pp * nm + j];
                            mm += op->B[pp * nm + i] * d->h_weights[pp] * op->B[pp * nm + j];
                        }
                        op->S[(size_t)i * nm + j] = k;
                        op->S[(size_t)nm * nm + (size_t)i * nm + j] = mm;
                    }
                    for (int pp = 0; pp < nq; ++pp) op->mvec[i] += op->B[pp * nm + i] * d->h_weights[pp];
                }
                // the separable kernels contract through the even-odd split: needs the point symmetry of a real basis
                double viol = 0.0, big = 0.0;
                for (int m2 = 0; m2 < 2; ++m2)
                    for (int i = 0; i < nm; ++i)
                        for (int j = 0; j < nm; ++j) {
                            const double a1 = op->S[(size_t)m2 * nm * nm + (size_t)i * nm + j], a2 = op->S[(size_t)m2 * nm * nm + (size_t)(nm - 1 - i) * nm + (nm - 1 - j)];
                            big = std::max(big, std::fabs(a1));
                            viol = std::max(viol, std::fabs(a1 - a2));
                        }
                if (!(viol <= 1e-11 * big)) { op->cartesian = false; op->S.clear(); op->mvec.clear(); }
            }
        }
    }
    if (d->d_cell_G && d->op_kind != B200FE_OP_LAPLACE && !op->cartesian)
        return fail(B200FE_ERR_UNSUPPORTED, "b200fe_op_create: mass / Helmholtz operators with on-the-fly geometry need axis-aligned cells");
    op_prepare_exclusive(*op);
    *out = reinterpret_cast<b200fe_op *>(op.release());
    return B200FE_OK;
}

void b200fe_op_destroy(b200fe_op *o)
{
    Operator *op = reinterpret_cast<Operator *>(o);
    if (!op) return;
    cg_release_work(op);
    delete op;  // ~Operator frees the owned device arrays and events
}

int b200fe_op_set_halo(b200fe_op *o, b200fe_halo *halo)
{
    B200FE_REQUIRE(o, "b200fe_op_set_halo: null operator");
    reinterpret_cast<Operator *>(o)->halo = reinterpret_cast<Halo *>(halo);
    return B200FE_OK;
}

int b200fe_op_set_constraints(b200fe_op *o, uint32_t n_rows, const uint32_t *h_hang_dof, const uint32_t *h_hang_row_ptr,
                              const uint32_t *h_hang_col, const double *h_hang_w)
{
    B200FE_REQUIRE(o, "b200fe_op_set_constraints: null operator");
    Operator &op = *reinterpret_cast<Operator *>(o);
    op.free_constraints();
    if (n_rows == 0) return B200FE_OK;
    B200FE_REQUIRE(h_hang_dof && h_hang_row_ptr, "b200fe_op_set_constraints: null pointer");
    const uint32_t nnz = h_hang_row_ptr[n_rows];
    B200FE_REQUIRE(h_hang_row_ptr[0] == 0, "b200fe_op_set_constraints: row_ptr[0] must be 0");
    B200FE_REQUIRE(nnz == 0 || (h_hang_col && h_hang_w), "b200fe_op_set_constraints: null pointer");
    // rows must be chain-free (a hanging DoF is never a parent) and inside the local vector
    std::vector<uint8_t> is_hanging(op.n_local(), 0);
    for (uint32_t r = 0; r < n_rows; ++r) {
        B200FE_REQUIRE(h_hang_dof[r] < op.n_local(), "b200fe_op_set_constraints: row %u constrains index %u outside the local vector", r, h_hang_dof[r]);
        B200FE_REQUIRE(h_hang_row_ptr[r] <= h_hang_row_ptr[r + 1], "b200fe_op_set_constraints: row_ptr not monotone at row %u", r);
        B200FE_REQUIRE(!is_hanging[h_hang_dof[r]], "b200fe_op_set_constraints: index %u constrained twice", h_hang_dof[r]);
        is_hanging[h_hang_dof[r]] = 1;
    }
    for (uint32_t k = 0; k < nnz; ++k) {
        B200FE_REQUIRE(h_hang_col[k] < op.n_local(), "b200fe_op_set_constraints: parent index %u outside the local vector", h_hang_col[k]);
        B200FE_REQUIRE(!is_hanging[h_hang_col[k]], "b200fe_op_set_constraints: constraint chain through index %u", h_hang_col[k]);
    }
    auto upload = [](auto **dst, const auto *src, size_t n) -> cudaError_t {
        cudaError_t e = cudaMalloc(dst, std::max<size_t>(n, 1) * sizeof(**dst));
        if (e == cudaSuccess && n) e = cudaMemcpy(*dst, src, n * sizeof(**dst), cudaMemcpyHostToDevice);
        return e;
    };
    cudaError_t e = upload(&op.d_hang_dof, h_hang_dof, n_rows);
    if (e == cudaSuccess) e = upload(&op.d_hang_ptr, h_hang_row_ptr, (size_t)n_rows + 1);
    if (e == cudaSuccess) e = upload(&op.d_hang_col, h_hang_col, nnz);
    if (e == cudaSuccess) e = upload(&op.d_hang_w, h_hang_w, nnz);
    if (e != cudaSuccess) {
        op.free_constraints();
        return fail_cuda(e, "b200fe_op_set_constraints");
    }
    op.n_hang = n_rows;
    op.h_hang_dof.assign(h_hang_dof, h_hang_dof + n_rows);
    op.h_hang_ptr.assign(h_hang_row_ptr, h_hang_row_ptr + n_rows + 1);
    op.h_hang_col.assign(h_hang_col, h_hang_col + nnz);
    op.h_hang_w.assign(h_hang_w, h_hang_w + nnz);
    return B200FE_OK;
}

int b200fe_op_set_face_constraints(b200fe_op *o, int p, uint32_t n_blocks, const uint32_t *h_face_parents,
                                   const uint32_t *h_face_children, const double *h_W)
{
    B200FE_REQUIRE(o, "b200fe_op_set_face_constraints: null operator");
    Operator &op = *reinterpret_cast<Operator *>(o);
    op.free_face_constraints();
    if (n_blocks == 0) return B200FE_OK;
    B200FE_REQUIRE(p == op.p, "b200fe_op_set_face_constraints: degree %d differs from the operator's %d", p, op.p);
    B200FE_REQUIRE(h_face_parents && h_face_children && h_W, "b200fe_op_set_face_constraints: null pointer");
    const int nm = p + 1, nf = 2 * p + 1;
    const size_t n_par = (size_t)n_blocks * nm * nm, n_chi = (size_t)n_blocks * nf * nf;
    std::vector<uint8_t> is_child(op.n_local(), 0);
    for (size_t i = 0; i < n_chi; ++i) {
        const uint32_t c = h_face_children[i];
        if (c == B200FE_INVALID_INDEX) continue;
        B200FE_REQUIRE(c < op.n_local(), "b200fe_op_set_face_constraints: child index %u outside the local vector", c);
        B200FE_REQUIRE(!is_child[c], "b200fe_op_set_face_constraints: index %u is the child of two blocks", c);
        is_child[c] = 1;
    }
    for (size_t i = 0; i < n_par; ++i) {
        const uint32_t q = h_face_parents[i];
        if (q == B200FE_INVALID_INDEX) continue;
        B200FE_REQUIRE(q < op.n_local(), "b200fe_op_set_face_constraints: parent index %u outside the local vector", q);
        B200FE_REQUIRE(!is_child[q], "b200fe_op_set_face_constraints: constraint chain through index %u", q);
    }
    cudaError_t e = cudaMalloc(&op.d_face_parents, n_par * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemcpy(op.d_face_parents, h_face_parents, n_par * sizeof(uint32_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&op.d_face_children, n_chi * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemcpy(op.d_face_children, h_face_children, n_chi * sizeof(uint32_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&op.d_face_W, (size_t)nf * nm * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(op.d_face_W, h_W, (size_t)nf * nm * sizeof(double), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        op.free_face_constraints();
        return fail_cuda(e, "b200fe_op_set_face_constraints");
    }
    op.n_face_blocks = n_blocks;
    op.face_nm = nm;
    op.face_nf = nf;
    return B200FE_OK;
}

int b200fe_op_distribute(b200fe_op *o, double *d_x, void *stream)
{
    B200FE_REQUIRE(o && d_x, "b200fe_op_distribute: null pointer");
    return op_distribute(*reinterpret_cast<Operator *>(o), d_x, false, (cudaStream_t)stream, 1);
}

int b200fe_op_vmult(b200fe_op *o, double *d_dst, const double *d_src, void *stream)
{
    B200FE_REQUIRE(o && d_dst && d_src, "b200fe_op_vmult: null pointer");
    B200FE_REQUIRE(d_dst != d_src, "b200fe_op_vmult: dst and src must not alias");
    return op_vmult(*reinterpret_cast<Operator *>(o), d_dst, d_src, nullptr, true, true, (cudaStream_t)stream);
}

int b200fe_op_vmult_components(b200fe_op *o, int n_components, double *d_dst, const double *d_src, void *stream)
{
    B200FE_REQUIRE(o && d_dst && d_src && n_components >= 1, "b200fe_op_vmult_components: bad arguments");
    B200FE_REQUIRE(d_dst != d_src, "b200fe_op_vmult_components: dst and src must not alias");
    return op_vmult(*reinterpret_cast<Operator *>(o), d_dst, d_src, nullptr, true, true, (cudaStream_t)stream, n_components);
}

int b200fe_op_vmult_dot(b200fe_op *o, double *d_dst, const double *d_src, double *d_dot, void *stream)
{
    B200FE_REQUIRE(o && d_dst && d_src && d_dot, "b200fe_op_vmult_dot: null pointer");
    B200FE_REQUIRE(d_dst != d_src, "b200fe_op_vmult_dot: dst and src must not alias");
    Operator &op = *reinterpret_cast<Operator *>(o);
    cudaStream_t s = (cudaStream_t)stream;
    B200FE_CUDA_TRY(cudaMemsetAsync(d_dot, 0, sizeof(double), s));
    if (int rc = op_vmult(op, d_dst, d_src, d_dot, true, true, s)) return rc;
    if (op.halo) return halo_allreduce_sum(*op.halo, d_dot, 1, s);
    return B200FE_OK;
}

int b200fe_op_vmult_dummy(b200fe_op *o, double *d_dst, const double *d_src, int ghost_exchange_on,
                          int computation_on, void *stream)
{
    B200FE_REQUIRE(o && d_dst && d_src, "b200fe_op_vmult_dummy: null pointer");
    return op_vmult(*reinterpret_cast<Operator *>(o), d_dst, d_src, nullptr, ghost_exchange_on != 0, computation_on != 0,
                    (cudaStream_t)stream);
}

int b200fe_op_diagonal(b200fe_op *o, double *d_diag, void *stream)
{
    B200FE_REQUIRE(o && d_diag, "b200fe_op_diagonal: null pointer");
    Operator &op = *reinterpret_cast<Operator *>(o);
    B200FE_REQUIRE(!(op.qop & QOP_LAPLACE) || op.d_G || op.cartesian, "b200fe_op_diagonal: needs the stored geometric factors (d_G) or axis-aligned cells");
    cudaStream_t s = (cudaStream_t)stream;
    if (op.cartesian && op.has_constraints())
        return fail(B200FE_ERR_UNSUPPORTED, "b200fe_op_diagonal: constrained operators need the stored geometric factors");
    if (op.has_constraints()) {
        if (op.n_hang == 0)
            return fail(B200FE_ERR_UNSUPPORTED, "b200fe_op_diagonal: an operator with face-structured constraints also needs the constraint rows "
                                                "(b200fe_op_set_constraints) for its diagonal");
        return op_diagonal_constrained(op, d_diag, s);
    }
    B200FE_CUDA_TRY(cudaMemsetAsync(d_diag, 0, sizeof(double) * op.n_local(), s));
    if (op.n_cells && op.cartesian) {
        const size_t n_entries = (size_t)op.n_cells * op.nm * op.nm * op.nm;
        cart_diagonal_kernel<<<(unsigned)std::min<size_t>((n_entries + 255) / 256, 148u * 16u), 256, 0, s>>>(cart_vecs(op), n_entries, op.nm, op.qop, op.d_cellG,
                                                                                                            op.d_idx, d_diag);
        B200FE_CUDA_TRY(cudaGetLastError());
    } else if (op.n_cells) {
        diagonal_kernel<<<std::min<uint32_t>(op.n_cells, 148u * 8u), 128, 0, s>>>(op.n_cells, op.nm, op.nq, op.qop, op.d_mats, op.d_G, op.d_JxW, op.d_idx, d_diag, nullptr);
        B200FE_CUDA_TRY(cudaGetLastError());
    }
    if (op.halo)
        if (int rc = halo_compress_add(*op.halo, d_diag, s)) return rc;
    if (op.n_constrained) {
        set_constrained_kernel<<<std::min<unsigned>((op.n_constrained + 255) / 256, 1184), 256, 0, s>>>(op.n_constrained, op.d_constrained, 1.0, d_diag);
        B200FE_CUDA_TRY(cudaGetLastError());
    }
    return B200FE_OK;
}

int b200fe_op_rhs_one(b200fe_op *o, double *d_b, void *stream)
{
    B200FE_REQUIRE(o && d_b, "b200fe_op_rhs_one: null pointer");
    Operator &op = *reinterpret_cast<Operator *>(o);
    B200FE_REQUIRE(op.d_JxW || op.cartesian || op.n_cells == 0, "b200fe_op_rhs_one: the operator was created without JxW");
    cudaStream_t s = (cudaStream_t)stream;
    B200FE_CUDA_TRY(cudaMemsetAsync(d_b, 0, sizeof(double) * op.n_local(), s));
    if (op.n_cells && op.cartesian && !op.d_JxW) {  // axis-aligned cells: int phi_i = det J * m_i m_j m_k
        const size_t n_entries = (size_t)op.n_cells * op.nm * op.nm * op.nm;
        cart_rhs_kernel<<<(unsigned)std::min<size_t>((n_entries + 255) / 256, 148u * 16u), 256, 0, s>>>(cart_vecs(op), n_entries, op.nm, op.d_cellG, op.d_idx, d_b);
        B200FE_CUDA_TRY(cudaGetLastError());
    } else if (op.n_cells) {
        rhs_one_kernel<<<std::min<uint32_t>(op.n_cells, 148u * 8u), 128, 0, s>>>(op.n_cells, op.nm, op.nq, op.d_mats, op.d_JxW, op.d_idx, d_b);
        B200FE_CUDA_TRY(cudaGetLastError());
    }
    if (int rc = op_condense(op, d_b, nullptr, s, 1)) return rc;  // b = C^T b_hat, hanging rows 0
    if (op.halo)
        if (int rc = halo_compress_add(*op.halo, d_b, s)) return rc;
    return B200FE_OK;
}

int b200fe_op_timing_enable(b200fe_op *o, int max_launches)
{
    B200FE_REQUIRE(o && max_launches >= 0, "b200fe_op_timing_enable: bad arguments");
    Operator &op = *reinterpret_cast<Operator *>(o);
    for (cudaEvent_t e : op.ev) cudaEventDestroy(e);
    op.ev.clear();
    op.ev_used = 0;
    op.timing = max_launches > 0;
    for (int i = 0; i < 2 * max_launches; ++i) {
        cudaEvent_t e;
        B200FE_CUDA_TRY(cudaEventCreate(&e));
        op.ev.push_back(e);
    }
    return B200FE_OK;
}

int b200fe_op_timing_read(b200fe_op *o, double *total_ms, int *launches)
{
    B200FE_REQUIRE(o && total_ms && launches, "b200fe_op_timing_read: null pointer");
    Operator &op = *reinterpret_cast<Operator *>(o);
    double sum = 0.0;
    for (size_t i = 0; i + 1 < op.ev_used; i += 2) {
        B200FE_CUDA_TRY(cudaEventSynchronize(op.ev[i + 1]));
        float ms = 0.f;
        B200FE_CUDA_TRY(cudaEventElapsedTime(&ms, op.ev[i], op.ev[i + 1]));
        sum += ms;
    }
    *total_ms = sum;
    *launches = (int)(op.ev_used / 2);
    op.ev_used = 0;
    return B200FE_OK;
}

unsigned long long b200fe_launch_count(void) { return g_launch_count; }

int b200fe_op_launch_info(b200fe_op *o, int *elems_per_block, int *num_blocks, int *threads_per_block,
                          int *smem_bytes, int *blocks_per_sm, int *regs_per_thread)
{
    B200FE_REQUIRE(o, "b200fe_op_launch_info: null operator");
    Operator &op = *reinterpret_cast<Operator *>(o);
    KArgs a{op.n_cells, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    LaunchInfo li{};
    B200FE_CUDA_TRY(op_launch_kernel(op, a, nullptr, &li, true));
    if (elems_per_block) *elems_per_block = li.elems_per_block;
    if (num_blocks) *num_blocks = li.num_blocks;
    if (threads_per_block) *threads_per_block = li.threads_per_block;
    if (smem_bytes) *smem_bytes = li.smem_bytes;
    if (blocks_per_sm) *blocks_per_sm = li.blocks_per_sm;
    if (regs_per_thread) *regs_per_thread = li.regs_per_thread;
    return B200FE_OK;
}

int b200fe_op_kernel_variant(b200fe_op *o, int *even_odd)
{
    B200FE_REQUIRE(o && even_odd, "b200fe_op_kernel_variant: null pointer");
    Operator &op = *reinterpret_cast<Operator *>(o);
    KArgs a{op.n_cells, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    LaunchInfo li{};
    B200FE_CUDA_TRY(op_launch_kernel(op, a, nullptr, &li, true));
    *even_odd = li.even_odd;
    return B200FE_OK;
}

int b200fe_op_cartesian(b200fe_op *o, int *on)
{
    Operator *op = reinterpret_cast<Operator *>(o);
    B200FE_REQUIRE(op && on, "b200fe_op_cartesian: null pointer");
    *on = op->cartesian ? 1 : 0;
    return B200FE_OK;
}

int b200fe_op_exclusive_interior(b200fe_op *o, int *on)
{
    Operator *op = reinterpret_cast<Operator *>(o);
    B200FE_REQUIRE(op && on, "b200fe_op_exclusive_interior: null pointer");
    *on = op->d_excl_mask != nullptr;
    return B200FE_OK;
}

}  // extern "C"
