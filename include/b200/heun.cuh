// Internal: the two update kernels of the Heun predictor/corrector.
//
// Reference: euler_step (solvers.cuh:114-125) and heun_step (:128-144). Both
// run in ORIGINAL cell order on the user-visible AoS arrays. The drift that is
// removed from x, y, z (mean force, or the force on a fixed point) is read
// from the device control block, where the pairwise sweep left it -- there is
// no device->host round trip between the sweep and the update.
//
// Rounding-relevant details kept from the reference:
//   predictor:  dX.xyz -= drift;  X1 = X + dX * dt
//   corrector:  dX1.xyz -= drift1; X += ((dX + dX1) * 0.5f) * dt;
//               old_v = (dX + dX1).xyz * 0.5
// The predictor does not write the drift-corrected dX back; the corrector
// redoes the same subtraction (bit-identical) instead of re-reading it.
#pragma once

#include <cuda_runtime.h>

#include "grid_build.cuh"
#include "layout.cuh"

namespace yb {

// X1 = X + (dX - drift) * dt, and -- for the grid solver -- the cube id of X1
// plus its bucket arrival rank (step 1 of the next grid build, fused). In a
// decomposed run (halo_flags != nullptr) also which faces of the brick X1 is
// close to, so that the halo round of the second stage need not read the
// positions again.
template<typename Pt, bool BIN>
__global__ void __launch_bounds__(256) predictor_step(
    const int* __restrict__ d_n, int n_max, float dt,
    const Pt* __restrict__ d_X, const Pt* __restrict__ d_dX,
    Pt* __restrict__ d_X1, Step_ctl* ctl, float cube_size, Grid_box box,
    int* __restrict__ key, int* __restrict__ arrival, int* count,
    Halo_faces faces = Halo_faces{}, unsigned char* __restrict__ halo_flags = nullptr)
{
    const int n = live_cells(d_n, n_max);
    const float fx = ctl->drift[0][0], fy = ctl->drift[0][1],
                fz = ctl->drift[0][2];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += gridDim.x * blockDim.x) {
        Pt dX = load_pt(d_dX, i);
        dX.x -= fx;
        dX.y -= fy;
        dX.z -= fz;
        const Pt X1 = load_pt(d_X, i) + dX * dt;
        store_pt(d_X1, i, X1);
        if (halo_flags) halo_flags[i] = halo_flags_of(X1.x, X1.y, X1.z, faces);
        if (BIN) {
            const int c = cube_of(
                X1.x, X1.y, X1.z, cube_size, box, &ctl->out_of_grid);
            key[i] = c;
            arrival[i] = atomicAdd(count + c, 1);
        }
    }
}

// X_out / v_out: where the new state goes if not in place (a decomposed run
// hands it straight to the migration pass, which re-stores it in d_X).
template<typename Pt>
__global__ void __launch_bounds__(256) corrector_step(
    const int* __restrict__ d_n, int n_max, float dt,
    const Pt* __restrict__ d_dX, const Pt* __restrict__ d_dX1, Pt* d_X,
    float3* __restrict__ d_old_v, const Step_ctl* __restrict__ ctl,
    Pt* X_out = nullptr, float3* v_out = nullptr)
{
    const int n = live_cells(d_n, n_max);
    const float fx0 = ctl->drift[0][0], fy0 = ctl->drift[0][1],
                fz0 = ctl->drift[0][2];
    const float fx1 = ctl->drift[1][0], fy1 = ctl->drift[1][1],
                fz1 = ctl->drift[1][2];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += gridDim.x * blockDim.x) {
        Pt dX = load_pt(d_dX, i);
        dX.x -= fx0;
        dX.y -= fy0;
        dX.z -= fz0;
        Pt dX1 = load_pt(d_dX1, i);
        dX1.x -= fx1;
        dX1.y -= fy1;
        dX1.z -= fz1;

        Pt X = load_pt_rw(d_X, i);
        X += (dX + dX1) * 0.5 * dt;
        store_pt(X_out ? X_out : d_X, i, X);

        float* v = reinterpret_cast<float*>((v_out ? v_out : d_old_v) + i);
        v[0] = (dX.x + dX1.x) * 0.5;
        v[1] = (dX.y + dX1.y) * 0.5;
        v[2] = (dX.z + dX1.z) * 0.5;
    }
}

}  // namespace yb
