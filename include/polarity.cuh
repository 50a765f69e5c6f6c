// Polarity forces in spherical coordinates.
//
// A polarity is a unit vector p given by its polar angle theta in [0, pi) and
// azimuth phi in [-pi, pi]. It lives either in a stand-alone Polarity or in
// two float members of a point type, selected by member pointers that default
// to &Pt::theta and &Pt::phi, so a point type may carry several polarities
// (e.g. epithelia_double_polarity.cu uses iota/chi as a second pair).
//
// All functions are usable on host and device. The formulas and -- because
// results must agree with the reference to rounding -- the order of the
// floating-point operations follow /root/reference/include/polarity.cuh:
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
    float3 term;
    term.x = -prod / dist * p.x + powf(prod, 2) / powf(dist, 2) * rx;
    term.y = -prod / dist * p.y + powf(prod, 2) / powf(dist, 2) * ry;
    term.z = -prod / dist * p.z + powf(prod, 2) / powf(dist, 2) * rz;
    return term;
}
}  // namespace yb_polarity


// Cartesian unit vector of a polarity.
template<typename Pt, float Pt::*theta = &Pt::theta, float Pt::*phi = &Pt::phi>
__device__ __host__ float3 pol_to_float3(Pt p)
{
    const float t = p.*theta;
    const float f = p.*phi;
    return float3{sinf(t) * cosf(f), sinf(t) * sinf(f), cosf(t)};
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
    return sinf(a.*theta) * sinf(p.theta) * cosf(a.*phi - p.phi) +
           cosf(a.*theta) * cosf(p.theta);
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
    dF.*theta = cosf(Xi.*theta) * sinf(p.theta) * cosf(Xi.*phi - p.phi) -
                sinf(Xi.*theta) * cosf(p.theta);
    const float sin_theta_i = sinf(Xi.*theta);
    if (fabs(sin_theta_i) > 1e-10)
        dF.*phi = -sinf(p.theta) * sinf(Xi.*phi - p.phi) / sin_theta_i;
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
    const float3 pi = pol_to_float3<Pt, theta, phi>(Xi);
    const float prodi = (pi.x * r.x + pi.y * r.y + pi.z * r.z) / dist;
    const Polarity r_hat = pt_to_pol(r, dist);

    // Angular part: turn p_i away from +-r_hat ...
    Pt dF = -prodi *
            unidirectional_polarization_force<Pt, theta, phi>(Xi, r_hat);

    // ... positional part from p_i ...
    const float3 from_i = yb_polarity::positional_bending_term(
        pi, prodi, dist, r.x, r.y, r.z);
    dF.x = from_i.x;
    dF.y = from_i.y;
    dF.z = from_i.z;

    // ... and from p_j = p_i - r's angles, via (p_j . r_ji / r)^2 / 2.
    const Polarity Xj{Xi.*theta - r.*theta, Xi.*phi - r.*phi};
    const float3 pj = pol_to_float3(Xj);
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
