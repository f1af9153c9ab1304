// Closed-form icosphere in meshzoo's vertex / cell order (util.py:17-50 -> meshzoo.icosa_sphere;
// layout from SURVEY.md Appendix B, parity unpinned: meshzoo is absent from the reference tree).
//
//   vertices: [12 corners | 30 edges x (k-1) | 20 faces x (k-1)(k-2)/2 interiors, row-major]
//   cells   : 20 faces x k^2 triangles; per face row i: (k-i) "up" then (k-i-1) "down"
//
// One thread per vertex / per triangle; nothing is read from memory except constant tables,
// so the d=2500 mesh (62.5 M vertices, 125 M triangles) is produced at store bandwidth on
// the device instead of 17 s on the host (nixis.py:242) plus a 1.5 GB upload.
//
// Vertex arithmetic is FP64 with every product/sum rounded once (no FMA) so that it is
// bit-identical to the numpy oracle (oracle/icosphere.py) -- see the formulas there.
#include "nxb_common.cuh"

#define PHI 1.618033988749895   /* (1 + sqrt(5)) / 2 rounded to double */

__constant__ double c_corner[12][3] = {
    {-1, PHI, 0}, {1, PHI, 0}, {-1, -PHI, 0}, {1, -PHI, 0},
    {0, -1, PHI}, {0, 1, PHI}, {0, -1, -PHI}, {0, 1, -PHI},
    {PHI, 0, -1}, {PHI, 0, 1}, {-PHI, 0, -1}, {-PHI, 0, 1}};

__constant__ int c_face[20][3] = {
    {0, 11, 5}, {0, 5, 1}, {0, 1, 7}, {0, 7, 10}, {0, 10, 11},
    {1, 5, 9}, {5, 11, 4}, {11, 10, 2}, {10, 7, 6}, {7, 1, 8},
    {3, 9, 4}, {3, 4, 2}, {3, 2, 6}, {3, 6, 8}, {3, 8, 9},
    {4, 9, 5}, {2, 4, 11}, {6, 2, 10}, {8, 6, 7}, {9, 8, 1}};

// CPython set-iteration order of the sorted corner pairs
__constant__ int c_edge[30][2] = {
    {3, 4}, {4, 9}, {8, 9}, {0, 5}, {2, 11}, {1, 9}, {0, 11}, {7, 10}, {6, 8}, {4, 5},
    {3, 9}, {3, 6}, {5, 9}, {4, 11}, {0, 1}, {0, 7}, {2, 4}, {10, 11}, {0, 10}, {1, 5},
    {2, 10}, {1, 8}, {6, 7}, {6, 10}, {3, 8}, {5, 11}, {2, 3}, {1, 7}, {2, 6}, {7, 8}};

// per face, per side (0: c0->c1, 1: c1->c2, 2: c2->c0): edge id * 2 + reversed
// (reversed = the side runs from the larger corner id to the smaller one; edge nodes are
// laid out from the smaller corner id).  Derived from c_face / c_edge; tests/test_mesh.py
// re-derives it.
__constant__ int c_face_edge[20][3] = {
    {12, 51, 7}, {6, 39, 29}, {28, 54, 31}, {30, 14, 37}, {36, 34, 13}, {38, 24, 11}, {50, 27, 18},
    {35, 41, 8}, {15, 45, 46}, {55, 42, 59}, {20, 3, 1}, {0, 33, 52}, {53, 56, 23}, {22, 16, 49},
    {48, 4, 21}, {2, 25, 19}, {32, 26, 9}, {57, 40, 47}, {17, 44, 58}, {5, 43, 10}};

// interiors before row r (r >= 1) of a face: sum_{q=1}^{r-1} (n-q-1)
__device__ __forceinline__ int64_t interior_before_row(int64_t n, int64_t r)
{
    return (r - 1) * (n - 1) - (r - 1) * r / 2;
}

// unit-sphere position of vertex v of the k-division icosphere, FP64, roundings as in oracle/icosphere.py
__device__ __forceinline__ void icosa_point(int64_t n, int64_t v, double &x, double &y, double &z)
{
    const int64_t n_edge = 30 * (n - 1), per_face = (n - 1) * (n - 2) / 2;
    if (v < 12) {
        x = c_corner[v][0]; y = c_corner[v][1]; z = c_corner[v][2];
    } else if (v < 12 + n_edge) {
        int64_t w = v - 12;
        int e = (int)(w / (n - 1));
        int64_t j = w % (n - 1) + 1;
        double t = __ddiv_rn((double)j, (double)n), u = __dsub_rn(1.0, t);
        const double *a = c_corner[c_edge[e][0]], *b = c_corner[c_edge[e][1]];
        x = __dadd_rn(__dmul_rn(u, a[0]), __dmul_rn(t, b[0]));
        y = __dadd_rn(__dmul_rn(u, a[1]), __dmul_rn(t, b[1]));
        z = __dadd_rn(__dmul_rn(u, a[2]), __dmul_rn(t, b[2]));
    } else {
        int64_t w = v - 12 - n_edge;
        int f = (int)(w / per_face);
        int64_t q = w % per_face;
        // row i (1..n-2): largest i with interior_before_row(i) <= q
        double disc = (double)(2 * n - 3) * (double)(2 * n - 3) - 8.0 * (double)q;
        int64_t i = (int64_t)(((double)(2 * n - 3) - sqrt(disc)) * 0.5) + 1;
        if (i < 1) i = 1;
        if (i > n - 2) i = n - 2;
        while (i > 1 && interior_before_row(n, i) > q) --i;
        while (i < n - 2 && interior_before_row(n, i + 1) <= q) ++i;
        int64_t j = q - interior_before_row(n, i) + 1;
        double bi = __ddiv_rn((double)i, (double)n), bj = __ddiv_rn((double)j, (double)n);
        double b0 = __dsub_rn(__dsub_rn(1.0, bi), bj);
        const double *c0 = c_corner[c_face[f][0]], *c1 = c_corner[c_face[f][1]], *c2 = c_corner[c_face[f][2]];
        x = __dadd_rn(__dadd_rn(__dmul_rn(b0, c0[0]), __dmul_rn(bj, c1[0])), __dmul_rn(bi, c2[0]));
        y = __dadd_rn(__dadd_rn(__dmul_rn(b0, c0[1]), __dmul_rn(bj, c1[1])), __dmul_rn(bi, c2[1]));
        z = __dadd_rn(__dadd_rn(__dmul_rn(b0, c0[2]), __dmul_rn(bj, c1[2])), __dmul_rn(bi, c2[2]));
    }
    double nrm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
    x = __ddiv_rn(x, nrm); y = __ddiv_rn(y, nrm); z = __ddiv_rn(z, nrm);
}

__global__ void __launch_bounds__(256)
icosa_points_kernel(int k, int64_t v_begin, int64_t v_end, float4 *__restrict__ out32, double *__restrict__ out64)
{
    for (int64_t v = v_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < v_end;
         v += (int64_t)gridDim.x * blockDim.x) {
        double x, y, z;
        icosa_point(k, v, x, y, z);
        int64_t o = v - v_begin;
        if (out32) out32[o] = make_float4((float)x, (float)y, (float)z, 0.0f);
        if (out64) { out64[3 * o] = x; out64[3 * o + 1] = y; out64[3 * o + 2] = z; }
    }
}

// Edge lengths of the radius-scaled icosphere for adjacency rows [v_begin, v_end) (GLOBAL vertex ids
// in adj_rows), positions recomputed in FP64 on the fly: dist[(v - v_begin) * 6 + q] =
// |R p_v - R p_n| (erosion.py:34-40, 227-229), rounded once to FP32; 0 for -1 pads.
__global__ void __launch_bounds__(256)
icosa_edge_lengths_kernel(int k, const int32_t *__restrict__ adj_rows, int64_t v_begin, int64_t v_end, double radius,
                          float *__restrict__ dist)
{
    for (int64_t v = v_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < v_end;
         v += (int64_t)gridDim.x * blockDim.x) {
        double x, y, z;
        icosa_point(k, v, x, y, z);
        x *= radius; y *= radius; z *= radius;
        const int64_t o = (v - v_begin) * 6;
        for (int q = 0; q < 6; ++q) {
            const int32_t n = adj_rows[o + q];
            float r = 0.0f;
            if (n >= 0) {
                double a, b, c;
                icosa_point(k, n, a, b, c);
                const double ax = x - a * radius, ay = y - b * radius, az = z - c * radius;
                r = (float)sqrt(ax * ax + ay * ay + az * az);
            }
            dist[o + q] = r;
        }
    }
}

NXB_API int nxb_mesh_icosa_edge_lengths(int k, const int32_t *adj_rows, int64_t v_begin, int64_t v_end, double radius,
                                        float *dist, void *stream)
{
    NXB_ARG(k >= 1 && k <= 14000 && adj_rows && dist);
    int64_t V = 10 * (int64_t)k * k + 2;
    NXB_ARG(0 <= v_begin && v_begin <= v_end && v_end <= V);
    if (v_end == v_begin) return NXB_OK;
    int grid = nxb_grid_resident(icosa_edge_lengths_kernel, 256, 0, (v_end - v_begin + 255) / 256);
    icosa_edge_lengths_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(k, adj_rows, v_begin, v_end, radius, dist);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

// local node (row r, col c) of face f -> global vertex id
__device__ __forceinline__ int32_t face_node(int64_t n, int f, int64_t r, int64_t c)
{
    const int64_t eb = 12, nm1 = n - 1;
    if (r == 0) {
        if (c == 0) return c_face[f][0];
        if (c == n) return c_face[f][1];
        int fe = c_face_edge[f][0];
        int64_t q = c - 1;
        return (int32_t)(eb + (fe >> 1) * nm1 + ((fe & 1) ? n - 2 - q : q));
    }
    if (r == n) return c_face[f][2];
    if (c == 0) {                       // side 2 runs c2->c0, i.e. downwards
        int fe = c_face_edge[f][2];
        int64_t q = r - 1;
        return (int32_t)(eb + (fe >> 1) * nm1 + ((fe & 1) ? q : n - 2 - q));
    }
    if (c == n - r) {
        int fe = c_face_edge[f][1];
        int64_t q = r - 1;
        return (int32_t)(eb + (fe >> 1) * nm1 + ((fe & 1) ? n - 2 - q : q));
    }
    int64_t base = 12 + 30 * nm1 + (int64_t)f * (nm1 * (n - 2) / 2);
    return (int32_t)(base + interior_before_row(n, r) + (c - 1));
}

__global__ void __launch_bounds__(256)
icosa_cells_kernel(int k, int64_t t_begin, int64_t t_end, int32_t *__restrict__ cells)
{
    const int64_t n = k, nn = n * n;
    for (int64_t t = t_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < t_end;
         t += (int64_t)gridDim.x * blockDim.x) {
        int f = (int)(t / nn);
        int64_t l = t % nn;
        // triangles before row i: 2*n*i - i*i ; row i = largest with that <= l
        int64_t i = (int64_t)((double)n - sqrt((double)(nn - l)));
        if (i < 0) i = 0;
        if (i > n - 1) i = n - 1;
        while (i > 0 && 2 * n * i - i * i > l) --i;
        while (i < n - 1 && 2 * n * (i + 1) - (i + 1) * (i + 1) <= l) ++i;
        int64_t r = l - (2 * n * i - i * i);
        int32_t a, b, c;
        if (r < n - i) {                 // up: (i,j) (i,j+1) (i+1,j)
            int64_t j = r;
            a = face_node(n, f, i, j); b = face_node(n, f, i, j + 1); c = face_node(n, f, i + 1, j);
        } else {                         // down: (i,j+1) (i+1,j+1) (i+1,j)
            int64_t j = r - (n - i);
            a = face_node(n, f, i, j + 1); b = face_node(n, f, i + 1, j + 1); c = face_node(n, f, i + 1, j);
        }
        int64_t o = 3 * (t - t_begin);
        cells[o] = a; cells[o + 1] = b; cells[o + 2] = c;
    }
}

NXB_API int nxb_mesh_icosa_points(int k, int64_t v_begin, int64_t v_end, nxb_float4 *xyz_f32, double *xyz_f64, void *stream)
{
    NXB_ARG(k >= 1 && k <= 14000);      // V and T must fit int32 vertex ids
    int64_t V = 10 * (int64_t)k * k + 2;
    NXB_ARG(0 <= v_begin && v_begin <= v_end && v_end <= V);
    if (v_end == v_begin) return NXB_OK;
    icosa_points_kernel<<<nxb_grid_resident(icosa_points_kernel, 256, 0, ((v_end - v_begin) + 256 - 1) / 256), 256, 0, (cudaStream_t)stream>>>(
        k, v_begin, v_end, (float4 *)xyz_f32, xyz_f64);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

NXB_API int nxb_mesh_icosa_cells(int k, int64_t t_begin, int64_t t_end, int32_t *cells, void *stream)
{
    NXB_ARG(k >= 1 && k <= 14000 && cells);
    int64_t T = 20 * (int64_t)k * k;
    NXB_ARG(0 <= t_begin && t_begin <= t_end && t_end <= T);
    if (t_end == t_begin) return NXB_OK;
    icosa_cells_kernel<<<nxb_grid_resident(icosa_cells_kernel, 256, 0, ((t_end - t_begin) + 256 - 1) / 256), 256, 0, (cudaStream_t)stream>>>(k, t_begin, t_end, cells);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

__global__ void __launch_bounds__(256)
xyz_f64_to_f32_kernel(const double *__restrict__ in, int64_t n, double scale, float4 *__restrict__ out)
{
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x)
        out[v] = make_float4((float)(in[3 * v] * scale), (float)(in[3 * v + 1] * scale), (float)(in[3 * v + 2] * scale), 0.0f);
}

NXB_API int nxb_xyz_f64_to_f32(const double *xyz_f64, int64_t n, double scale, nxb_float4 *xyz_f32, void *stream)
{
    NXB_ARG(n >= 0);
    if (n == 0) return NXB_OK;
    NXB_ARG(xyz_f64 && xyz_f32);
    xyz_f64_to_f32_kernel<<<nxb_grid_resident(xyz_f64_to_f32_kernel, 256, 0, ((n) + 256 - 1) / 256), 256, 0, (cudaStream_t)stream>>>(xyz_f64, n, scale, (float4 *)xyz_f32);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}
