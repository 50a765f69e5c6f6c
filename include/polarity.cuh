// Polarity forces in spherical coordinates.
//
// A polarity is a unit vector p given by its polar angle theta in [0, pi) and
// azimuth phi in [-pi, pi]. It lives either in a stand-alone Polarity or in
// two float members of a point type, selected by member pointers that default
// to &Pt::theta and &Pt::phi, so a point type may carry several polarities
// (e.g. epithelia_double_polarity.cu uses iota/chi as a second pair).
//
// All functions are usable on host and device. The formulas follow
// /root/reference/include/polarity.cuh; where the same values are needed the
// order of the floating-point operations is kept, so results agree with the
// reference to rounding. Two things are done differently because these
// functions are the inner loop of every epithelial model:
//  * sine and cosine of one angle come from one sincosf (one range reduction);
//  * bending_force never converts r to angles and back. The reference computes
//    r_hat = (acosf(r.z / d), atan2(r.y, r.x)) and then sin/cos of those angles;
//    here sin(theta_r) cos(phi_i - phi_r) = (cos(phi_i) r.x + sin(phi_i) r.y) / d
//    etc. are used directly, which is the same quantity with fewer roundings
//    (no acosf/atan2f, 4 instead of 14 trigonometric evaluations per pair).
//    Differences to the reference's own build are a few ulp of the force;
//    tests/test_gpu_parity.py bounds them against golden vectors of that build.
// Reference lines:
//   pol_to_float3 :13-21, pt_to_pol :23-39, pol_dot_product :41-46,
//   unidirectional_polarization_force :48-60, bidirectional_… :62-69,
//   bending_force :71-94, apical_constriction_force :96-122,
//   orthonormal :125-131, migration_force :133-164.
#pragma once

#include <math.h>

#include "utils.cuh"


struct Polarity {
    float theta, phi;
};

namespace yb_polarity {
struct Sin_cos {
    float sin, cos;
};

__device__ __host__ inline Sin_cos sin_cos(float angle)
{
    Sin_cos result;
    sincosf(angle, &result.sin, &result.cos);
    return result;
}

// |sin(theta)| below which bending_force takes the reference's exact route.
constexpr float pole_guard = 0.1f;

// The pair (theta, phi) of a point type as a Polarity value.
template<typename Pt, float Pt::*theta, float Pt::*phi>
__device__ __host__ inline Polarity angles_of(const Pt& X)
{
    return Polarity{X.*theta, X.*phi};
}

// Derivative of U = (p . r_hat)^2 / 2 with respect to the position of the
// cell that owns p:  -(p.r_hat)/d * p  +  (p.r_hat)^2/d^2 * r.
__device__ __host__ inline float3 positional_bending_term(
    float3 p, float prod, float dist, float rx, float ry, float rz)
{
    // powf(x, 2) of the reference is x * x up to its rounding
    const float along_p = -prod / dist;
    const float along_r = (prod * prod) / (dist * dist);
    float3 term;
    term.x = along_p * p.x + along_r * rx;
    term.y = along_p * p.y + along_r * ry;
    term.z = along_p * p.z + along_r * rz;
    return term;
}
}  // namespace yb_polarity


// Cartesian unit vector of a polarity.
template<typename Pt, float Pt::*theta = &Pt::theta, float Pt::*phi = &Pt::phi>
__device__ __host__ float3 pol_to_float3(Pt p)
{
    const yb_polarity::Sin_cos t = yb_polarity::sin_cos(p.*theta);
    const yb_polarity::Sin_cos f = yb_polarity::sin_cos(p.*phi);
    return float3{t.sin * f.cos, t.sin * f.sin, t.cos};
}

// Direction of r as a polarity; dist must be |r|.
template<typename Pt>
__device__ __host__ Polarity pt_to_pol(Pt r, float dist)
{
    return Polarity{acosf(r.z / dist), atan2(r.y, r.x)};
}

template<typename Pt>
__device__ __host__ Polarity pt_to_pol(Pt r)
{
#ifdef __CUDA_ARCH__
    const float dist = norm3df(r.x, r.y, r.z);
#else
    const float dist = sqrt(r.x * r.x + r.y * r.y + r.z * r.z);
#endif
    return pt_to_pol(r, dist);
}

// p_a . p from the spherical law of cosines.
template<typename Pt, float Pt::*theta = &Pt::theta, float Pt::*phi = &Pt::phi>
__device__ __host__ float pol_dot_product(Pt a, Polarity p)
{
    const yb_polarity::Sin_cos ta = yb_polarity::sin_cos(a.*theta);
    const yb_polarity::Sin_cos tp = yb_polarity::sin_cos(p.theta);
    return ta.sin * tp.sin * cosf(a.*phi - p.phi) + ta.cos * tp.cos;
}

// Same, with the second polarity taken from another point (not part of the
// reference API; lets pol_dot_product(p, cells.h_X[i]) compile).
template<typename Pt_a, typename Pt_b,
    typename = decltype(Pt_b::x)>
__device__ __host__ float pol_dot_product(Pt_a a, Pt_b b)
{
    return pol_dot_product(a, Polarity{b.theta, b.phi});
}


// Torque aligning Xi's polarity WITH p, from U = -sum(p_i . p_j): the gradient
// of p_i . p in (theta, phi), the phi component divided by sin(theta)^2's
// metric factor. Close to the poles the azimuthal part is dropped.
template<typename Pt, float Pt::*theta = &Pt::theta, float Pt::*phi = &Pt::phi>
__device__ __host__ Pt unidirectional_polarization_force(Pt Xi, Polarity p)
{
    Pt dF{0};
    const yb_polarity::Sin_cos ti = yb_polarity::sin_cos(Xi.*theta);
    const yb_polarity::Sin_cos tp = yb_polarity::sin_cos(p.theta);
    const yb_polarity::Sin_cos df = yb_polarity::sin_cos(Xi.*phi - p.phi);
    dF.*theta = ti.cos * tp.sin * df.cos - ti.sin * tp.cos;
    if (fabs(ti.sin) > 1e-10) dF.*phi = -tp.sin * df.sin / ti.sin;
    return dF;
}

// Torque aligning Xi's polarity with p OR -p ("planar cell polarity"), from
// U = -sum (p_i . p_j)^2 / 2.
template<typename Pt, float Pt::*theta = &Pt::theta, float Pt::*phi = &Pt::phi>
__device__ __host__ Pt bidirectional_polarization_force(Pt Xi, Polarity p)
{
    const float alignment = pol_dot_product<Pt, theta, phi>(Xi, p);
    return alignment *
           unidirectional_polarization_force<Pt, theta, phi>(Xi, p);
}

// Convenience overloads taking the partner's polarity from a point of the same
// type (the upstream tests and examples/polarization.cu call it this way).
template<typename Pt, float Pt::*theta = &Pt::theta, float Pt::*phi = &Pt::phi>
__device__ __host__ Pt unidirectional_polarization_force(Pt Xi, Pt Xj)
{
    return unidirectional_polarization_force<Pt, theta, phi>(
        Xi, yb_polarity::angles_of<Pt, theta, phi>(Xj));
}

template<typename Pt, float Pt::*theta = &Pt::theta, float Pt::*phi = &Pt::phi>
__device__ __host__ Pt bidirectional_polarization_force(Pt Xi, Pt Xj)
{
    return bidirectional_polarization_force<Pt, theta, phi>(
        Xi, yb_polarity::angles_of<Pt, theta, phi>(Xj));
}


// Resistance of an epithelial sheet against bending, from
// U = sum (p_i . r_ij / |r_ij|)^2 / 2: polarities want to stand normal to the
// connections to their neighbours. r = Xi - Xj (all members), dist = |r|.
// Returns the torque on p_i and the force on i from both (p_i . r_hat)^2 / 2
// and (p_j . r_hat)^2 / 2.
template<typename Pt, float Pt::*theta = &Pt::theta, float Pt::*phi = &Pt::phi>
__device__ __host__ Pt bending_force(Pt Xi, Pt r, float dist)
{
    const yb_polarity::Sin_cos ti = yb_polarity::sin_cos(Xi.*theta);
    const yb_polarity::Sin_cos fi = yb_polarity::sin_cos(Xi.*phi);
    const float3 pi{ti.sin * fi.cos, ti.sin * fi.sin, ti.cos};
    const float prodi = (pi.x * r.x + pi.y * r.y + pi.z * r.z) / dist;

    // Angular part: -prodi * d(p_i . r_hat)/d(theta_i, phi_i). With r_hat's
    // angles (theta_r, phi_r):  sin(theta_r) cos(phi_i - phi_r) = along / dist,
    // sin(theta_r) sin(phi_i - phi_r) = across / dist, cos(theta_r) = r.z / dist.
    Pt dF{0};
    if (fabs(ti.sin) > yb_polarity::pole_guard) {
        const float along = fi.cos * r.x + fi.sin * r.y;
        const float across = fi.sin * r.x - fi.cos * r.y;
        dF.*theta = -prodi * ((ti.cos * along - ti.sin * r.z) / dist);
        dF.*phi = -prodi * (-(across / dist) / ti.sin);
    } else {
        // Close to the poles of the coordinate system d(phi)/dt ~ 1 / sin(theta)
        // amplifies every rounding difference, so there the torque is computed
        // exactly the way the reference does it, through r_hat's angles.
        const Polarity r_hat = pt_to_pol(r, dist);
        const yb_polarity::Sin_cos tr = yb_polarity::sin_cos(r_hat.theta);
        const yb_polarity::Sin_cos df = yb_polarity::sin_cos(Xi.*phi - r_hat.phi);
        dF.*theta = -prodi * (ti.cos * tr.sin * df.cos - ti.sin * tr.cos);
        if (fabs(ti.sin) > 1e-10)
            dF.*phi = -prodi * (-tr.sin * df.sin / ti.sin);
    }

    // ... positional part from p_i ...
    const float3 from_i = yb_polarity::positional_bending_term(
        pi, prodi, dist, r.x, r.y, r.z);
    dF.x = from_i.x;
    dF.y = from_i.y;
    dF.z = from_i.z;

    // ... and from p_j = p_i - r's angles, via (p_j . r_ji / r)^2 / 2.
    const yb_polarity::Sin_cos tj = yb_polarity::sin_cos(Xi.*theta - r.*theta);
    const yb_polarity::Sin_cos fj = yb_polarity::sin_cos(Xi.*phi - r.*phi);
    const float3 pj{tj.sin * fj.cos, tj.sin * fj.sin, tj.cos};
    const float prodj = (pj.x * r.x + pj.y * r.y + pj.z * r.z) / dist;
    const float3 from_j = yb_polarity::positional_bending_term(
        pj, prodj, dist, r.x, r.y, r.z);
    dF.x += from_j.x;
    dF.y += from_j.y;
    dF.z += from_j.z;

    return dF;
}

// Bending force with a preferred angle between p_i and r_ij other than 90
// degrees, i.e. wedge-shaped cells; pref_angle = pi/2 gives bending_force.
template<typename Pt>
__device__ __host__ Pt apical_constriction_force(
    Pt Xi, Pt r, float dist, float pref_angle)
{
    const float3 pi = pol_to_float3(Xi);
    const float prodi =
        (pi.x * r.x + pi.y * r.y + pi.z * r.z) / dist + cosf(pref_angle);
    const Polarity r_hat = pt_to_pol(r, dist);

    Pt dF = -prodi * unidirectional_polarization_force(Xi, r_hat);

    const float3 from_i = yb_polarity::positional_bending_term(
        pi, prodi, dist, r.x, r.y, r.z);
    dF.x = from_i.x;
    dF.y = from_i.y;
    dF.z = from_i.z;

    const Polarity Xj{Xi.theta - r.theta, Xi.phi - r.phi};
    const float3 pj = pol_to_float3(Xj);
    const float prodj =
        (pj.x * r.x + pj.y * r.y + pj.z * r.z) / dist - cosf(pref_angle);
    const float3 from_j = yb_polarity::positional_bending_term(
        pj, prodj, dist, r.x, r.y, r.z);
    dF.x += from_j.x;
    dF.y += from_j.y;
    dF.z += from_j.z;

    return dF;
}


// Unit vector in the plane of r and p that is perpendicular to p.
template<typename Pt>
__device__ __host__ float3 orthonormal(Pt r, float3 p)
{
    const float3 r3{r.x, r.y, r.z};
    const float3 rejected = r3 - dot_product(r3, p) * p;
    return rejected / sqrt(dot_product(rejected, rejected));
}

// Mono-polar migration (https://doi.org/10.1016/B978-0-12-405926-9.00016-2):
// a cell whose polarity points towards neighbour j crawls around it, and is
// pushed aside by neighbours that crawl towards it.
template<typename Pt, float Pt::*theta = &Pt::theta, float Pt::*phi = &Pt::phi>
__device__ __host__ Pt migration_force(Pt Xi, Pt r, float dist)
{
    Pt dF{0};
    const Polarity r_hat = pt_to_pol(r, dist);

    // i pulls itself around j
    if ((Xi.phi != 0) or (Xi.theta != 0)) {
        if (pol_dot_product<Pt, theta, phi>(Xi, r_hat) <= -0.15) {
            const float3 pi = pol_to_float3<Pt, theta, phi>(Xi);
            const float3 pi_T = orthonormal(r, pi);
            dF.x = 0.6 * pi.x + 0.8 * pi_T.x;
            dF.y = 0.6 * pi.y + 0.8 * pi_T.y;
            dF.z = 0.6 * pi.z + 0.8 * pi_T.z;
        }
    }

    // i gets pushed aside by j
    const Polarity Xj{Xi.*theta - r.*theta, Xi.*phi - r.*phi};
    if ((Xj.phi > 1e-10) or (Xj.theta > 1e-10)) {
        if (pol_dot_product(Xj, r_hat) >= 0.15) {
            const float3 pj = pol_to_float3(Xj);
            const float3 pj_T = orthonormal(-r, pj);
            dF.x -= 0.6 * pj.x + 0.8 * pj_T.x;
            dF.y -= 0.6 * pj.y + 0.8 * pj_T.y;
            dF.z -= 0.6 * pj.z + 0.8 * pj_T.z;
        }
    }

    return dF;
}
