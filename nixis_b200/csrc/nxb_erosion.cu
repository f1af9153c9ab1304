// Erosion sweeps over the ELL neighbour table -- erosion.py:34-40, 76-99, 197-279.
//
// HBM-bound stencil.  Algorithmic traffic per vertex-iteration of iteration3:
//   own h,w,s read 12 B + h,w,s write 12 B + adjacency row 24 B + own xyz 12 B = 60 B;
// neighbour values are some other vertex's compulsory read and come from L1/L2: the meshzoo
// order is row-major inside each icosahedron face, so the 6 neighbours of vertex v are
// v+-1 and two short runs one mesh row above / below.
//
// Ping-pong buffers replace the reference's three np.copy + copy-back pass (erosion.py:199-201,
// 274-277); `water += rain` (erosion.py:182-183) is fused into the reads.
#include "nxb_common.cuh"

__device__ __forceinline__ void load_adj_row(const int32_t *__restrict__ adj, int64_t v, int32_t (&row)[6])
{
    const int2 *p = reinterpret_cast<const int2 *>(adj + v * 6);
    int2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    row[0] = a.x; row[1] = a.y; row[2] = b.x; row[3] = b.y; row[4] = c.x; row[5] = c.y;
}

__global__ void __launch_bounds__(256)
erode3_kernel(const float4 *__restrict__ xyz, const int32_t *__restrict__ adj,
              const float *__restrict__ h_in, const float *__restrict__ w_in, const float *__restrict__ s_in,
              float *__restrict__ h_out, float *__restrict__ w_out, float *__restrict__ s_out,
              int64_t v_begin, int64_t v_end, float rain, float radius)
{
    const float evaporation = (float)(0.1 / 320), solubility = (float)(0.01 / 320), capacity = (float)(0.2 / 320);
    for (int64_t i = v_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < v_end;
         i += (int64_t)gridDim.x * blockDim.x) {
        int32_t row[6];
        load_adj_row(adj, i, row);
        const float4 pi = __ldg(xyz + i);
        const float me = h_in[i];
        const float wat_i = w_in[i] + rain;
        const float sed_i = s_in[i];
        float sed_amt = sed_i, wat_amt = wat_i;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const int32_t n = row[q];
            if (n < 0) continue;
            const float4 pn = __ldg(xyz + n);
            const float hn = __ldg(h_in + n);
            const float wn = __ldg(w_in + n) + rain;
            const float ax = pi.x - pn.x, ay = pi.y - pn.y, az = pi.z - pn.z;
            const float d = radius * sqrtf(ax * ax + ay * ay + az * az);
            // slope = (hn - me) / (d + 1e-5): only its sign is used and d + 1e-5 > 0
            const float dh = hn - me;
            if (dh > 0.0f)      { sed_amt += solubility * wn; wat_amt += wn * d; }
            else if (dh < 0.0f) { sed_amt -= solubility * wn; wat_amt -= wn * d; }
        }
        float hh = me - sed_amt;
        float ss = sed_i + sed_amt;
        float ww = wat_i + (wat_amt - wat_amt * evaporation);
        const float cw = capacity * ww;
        if (ss > cw) { hh += ss - cw; ss -= ss - cw; }
        h_out[i] = hh; w_out[i] = ww; s_out[i] = ss;
    }
}

NXB_API int nxb_erode3_step_f32(const nxb_float4 *xyz_unit, const int32_t *adj,
                                const float *h_in, const float *w_in, const float *s_in,
                                float *h_out, float *w_out, float *s_out,
                                int64_t v_begin, int64_t v_end, float rain, float radius, void *stream)
{
    NXB_ARG(v_begin >= 0 && v_end >= v_begin);
    if (v_end == v_begin) return NXB_OK;
    NXB_ARG(xyz_unit && adj && h_in && w_in && s_in && h_out && w_out && s_out);
    NXB_ARG(h_in != h_out && w_in != w_out && s_in != s_out);
    erode3_kernel<<<nxb_grid_for(v_end - v_begin, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        (const float4 *)xyz_unit, adj, h_in, w_in, s_in, h_out, w_out, s_out, v_begin, v_end, rain, radius);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

// erosion.py:76-99
__global__ void __launch_bounds__(256)
erode1_kernel(const int32_t *__restrict__ adj, const float *__restrict__ h_in, float *__restrict__ h_out,
              int64_t v_begin, int64_t v_end)
{
    for (int64_t i = v_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < v_end;
         i += (int64_t)gridDim.x * blockDim.x) {
        int32_t row[6];
        load_adj_row(adj, i, row);
        const float me = h_in[i];
        float amt = 0.0f;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const int32_t n = row[q];
            if (n < 0) continue;
            const float hn = __ldg(h_in + n);
            if (hn > me) amt += 0.0005f;
            else if (hn < me) amt -= 0.0005f;
        }
        h_out[i] = me + amt;
    }
}

NXB_API int nxb_erode1_step_f32(const int32_t *adj, const float *h_in, float *h_out,
                                int64_t v_begin, int64_t v_end, void *stream)
{
    NXB_ARG(v_begin >= 0 && v_end >= v_begin);
    if (v_end == v_begin) return NXB_OK;
    NXB_ARG(adj && h_in && h_out && h_in != h_out);
    erode1_kernel<<<nxb_grid_for(v_end - v_begin, 256, 8), 256, 0, (cudaStream_t)stream>>>(adj, h_in, h_out, v_begin, v_end);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}
