/*
 * kdsl_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the variational-Monte-Carlo sampling path of
 * hz-xiaxz/KagomeDSL.jl (pure Julia; cannot be executed in this image, see
 * DESIGN.md "Oracle").  Every function cites the reference file:line it
 * follows (paths relative to /root/reference).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product (libkdsl.so) never links or calls it.
 *
 * Parity status: the reference's own unit-test known-answer vectors
 * (test/test-Hamiltonian.jl, test/test-MonteCarlo.jl, test/test-Lattice.jl)
 * are checked in tests/test_oracle_golden.py.  The reference tests pin no
 * sweep trajectory, RNG stream, or energy value, so for those the oracle is
 * "parity unpinned" (anchored on exact small-lattice enumeration instead).
 *
 * Two scalar instantiations of the MC core are generated from
 * oracle_mc_core.inc: c128 (ComplexF64, the reference's storage type,
 * src/Hamiltonian.jl:420-427) and f64 (real, valid for B = 0).
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* Lattice geometry: src/Lattice.jl:68-97                              */
/* ------------------------------------------------------------------ */
typedef struct {
    int n1, n2;          /* src/Lattice.jl:19-20 */
    double t;            /* :22 */
    double a1[2], a2[2]; /* :24-25 */
    double r[6][2];      /* :27 */
    int pbc[2], anti[2]; /* :29-30 */
} ko_lattice;

/* src/Lattice.jl:68-92 (constructor) + :3-15 (validate_boundary_conditions).
 * returns 0 ok, -1 antiPBC without PBC (ArgumentError), -2 n1 odd (AssertionError) */
int ko_lattice_init(ko_lattice *lat, double t, int n1, int n2, int pbc1, int pbc2,
                    int anti1, int anti2) {
    if ((anti1 && !pbc1) || (anti2 && !pbc2)) return -1;
    if (n1 % 2 != 0) return -2;
    double a = 2.0 * t;
    lat->n1 = n1; lat->n2 = n2; lat->t = t;
    lat->a1[0] = 2.0 * a;           lat->a1[1] = 0.0;
    lat->a2[0] = 0.5 * a;           lat->a2[1] = 0.5 * sqrt(3.0) * a;
    lat->r[0][0] = 0.0;             lat->r[0][1] = 0.0;
    lat->r[1][0] = 0.25 * lat->a1[0]; lat->r[1][1] = 0.25 * lat->a1[1];
    lat->r[2][0] = 0.5 * lat->a2[0];  lat->r[2][1] = 0.5 * lat->a2[1];
    lat->r[3][0] = 0.5 * lat->a1[0];  lat->r[3][1] = 0.5 * lat->a1[1];
    lat->r[4][0] = 0.75 * lat->a1[0]; lat->r[4][1] = 0.75 * lat->a1[1];
    lat->r[5][0] = 2.5 * t;         lat->r[5][1] = 0.5 * sqrt(3.0) * t;
    lat->pbc[0] = pbc1; lat->pbc[1] = pbc2; lat->anti[0] = anti1; lat->anti[1] = anti2;
    return 0;
}

/* src/Lattice.jl:97 */
int ko_ns(const ko_lattice *lat) { return lat->n1 * lat->n2 * 3; }

/* src/Hamiltonian.jl:28-38 unitcell_coord (s is 1-based). returns -1 on the @assert */
int ko_unitcell_coord(const ko_lattice *lat, int s, double out[2]) {
    int n1 = lat->n1 / 2, n2 = lat->n2;
    int ns = n1 * n2 * 6;
    if (!(s >= 1 && s <= ns)) return -1;
    int uc = (s - 1) / 6;
    out[0] = (uc % n1) * lat->a1[0] + (uc / n1) * lat->a2[0];
    out[1] = (uc % n1) * lat->a1[1] + (uc / n1) * lat->a2[1];
    return 0;
}

/* src/Hamiltonian.jl:56-76 unitcell_diff */
void ko_unitcell_diff(const ko_lattice *lat, const double c1[2], const double c2[2],
                      int *dx, int *dy) {
    double d0 = c1[0] - c2[0], d1 = c1[1] - c2[1];
    const double *a1 = lat->a1, *a2 = lat->a2;
    double det = a1[0] * a2[1] - a1[1] * a2[0];
    *dx = (int)rint((a2[1] * d0 - a2[0] * d1) / det);   /* round(Int, .) = ties-to-even */
    *dy = (int)rint((-a1[1] * d0 + a1[0] * d1) / det);
}

/* src/Hamiltonian.jl:279-283 get_site_coord */
int ko_get_site_coord(const ko_lattice *lat, int s, double out[2]) {
    int label = (s - 1) % 6;
    double uc[2];
    if (ko_unitcell_coord(lat, s, uc)) return -1;
    out[0] = uc[0] + lat->r[label][0];
    out[1] = uc[1] + lat->r[label][1];
    return 0;
}

/* src/Hamiltonian.jl:99-153 get_boundary_shifts.
 * out: up to 9 triples (dx, dy, sign). returns count, or -1 on an @assert. */
int ko_get_boundary_shifts(const ko_lattice *lat, int s1, int s2, int out_dx[9],
                           int out_dy[9], double out_sign[9]) {
    if (s1 == s2) return -1;
    int n1 = lat->n1 / 2, n2 = lat->n2;
    int ns = n1 * n2 * 6;
    if (!(s1 >= 1 && s1 <= ns) || !(s2 >= 1 && s2 <= ns)) return -1;
    double u1[2], u2[2];
    ko_unitcell_coord(lat, s1, u1);
    ko_unitcell_coord(lat, s2, u2);
    int dx, dy;
    ko_unitcell_diff(lat, u2, u1, &dx, &dy);
    int n = 0;
    out_dx[n] = dx; out_dy[n] = dy; out_sign[n] = 1.0; n++;
    if (!lat->pbc[0] && !lat->pbc[1]) return n;         /* :91-93 */
    int sx[3], sy[3], nx, ny;
    if (lat->pbc[0]) { sx[0] = -n1; sx[1] = 0; sx[2] = n1; nx = 3; } else { sx[0] = 0; nx = 1; }
    if (lat->pbc[1]) { sy[0] = -n2; sy[1] = 0; sy[2] = n2; ny = 3; } else { sy[0] = 0; ny = 1; }
    for (int ix = 0; ix < nx; ix++)
        for (int iy = 0; iy < ny; iy++) {
            int sh1 = sx[ix], sh2 = sy[iy];
            if (sh1 == 0 && sh2 == 0) continue;           /* :100 (shifts is never empty) */
            double sign = 1.0;
            if (lat->anti[0] && sh1 != 0) sign *= -1.0;
            if (lat->anti[1] && sh2 != 0) sign *= -1.0;
            /* :122 unique(shifts): drop exact duplicates (possible only when n1 or n2 is 0) */
            int dup = 0;
            for (int q = 0; q < n; q++)
                if (out_dx[q] == dx + sh1 && out_dy[q] == dy + sh2 && out_sign[q] == sign) dup = 1;
            if (dup) continue;
            out_dx[n] = dx + sh1; out_dy[n] = dy + sh2; out_sign[n] = sign; n++;
        }
    return n;
}

/* link tables are passed as flat int arrays:
 *   link_in   [n_in][3]    = (label1, label2, value)          src/Hamiltonian.jl:219-240
 *   link_inter[n_inter][5] = (label1, label2, dx, dy, value)  src/Hamiltonian.jl:250-261 */
static int find_link_in(const int *t, int n, int l1, int l2, int *val) {
    for (int q = 0; q < n; q++)
        if (t[3 * q] == l1 && t[3 * q + 1] == l2) { *val = t[3 * q + 2]; return 1; }
    return 0;
}
static int find_link_inter(const int *t, int n, int l1, int l2, int dx, int dy, int *val) {
    for (int q = 0; q < n; q++)
        if (t[5 * q] == l1 && t[5 * q + 1] == l2 && t[5 * q + 2] == dx && t[5 * q + 3] == dy) {
            *val = t[5 * q + 4];
            return 1;
        }
    return 0;
}

/* src/Hamiltonian.jl:177-207 apply_boundary_conditions!  (tunneling col-major ns x ns)
 * returns 0, or -1 on an @assert (same cell / bad index) */
int ko_apply_boundary_conditions(double _Complex *tunneling, int ld, const ko_lattice *lat,
                                 int s1, int s2, const int *link_inter, int n_inter,
                                 double B) {
    int n1 = lat->n1 / 2, n2 = lat->n2;
    int nsl = n1 * n2 * 6;
    if (!(s1 >= 1 && s1 <= nsl) || !(s2 >= 1 && s2 <= nsl)) return -1;
    int cell1 = (s1 - 1) / 6 + 1, cell2 = (s2 - 1) / 6 + 1;
    if (cell1 == cell2) return -1;
    int label1 = (s1 - 1) % 6 + 1, label2 = (s2 - 1) % 6 + 1;
    int sdx[9], sdy[9];
    double ssg[9];
    int nsh = ko_get_boundary_shifts(lat, s1, s2, sdx, sdy, ssg);
    if (nsh < 0) return -1;
    double r1[2], ruc1[2], r2[2], ruc2[2], dr2[2];
    ko_get_site_coord(lat, s1, r1);
    ko_unitcell_coord(lat, s1, ruc1);
    ko_get_site_coord(lat, s2, r2);
    ko_unitcell_coord(lat, s2, ruc2);
    dr2[0] = r2[0] - ruc2[0];
    dr2[1] = r2[1] - ruc2[1];
    for (int q = 0; q < nsh; q++) {
        int val;
        if (find_link_inter(link_inter, n_inter, label1, label2, sdx[q], sdy[q], &val)) {
            double x2 = ruc1[0] + sdx[q] * lat->a1[0] + sdy[q] * lat->a2[0] + dr2[0];
            double y2 = ruc1[1] + sdx[q] * lat->a1[1] + sdy[q] * lat->a2[1] + dr2[1];
            double phase = (B / 2.0) * (r1[0] + x2) * (y2 - r1[1]);
            double _Complex hop = cexp(I * phase);
            tunneling[(size_t)(s2 - 1) * ld + (s1 - 1)] += ssg[q] * (double)val * hop;
        }
    }
    return 0;
}

/* src/Hamiltonian.jl:308-354 Hmat.  H col-major ns x ns (complex).
 * returns 0 ok, -3 "tunneling matrix must be upper triangular" */
int ko_hmat(const ko_lattice *lat, const int *link_in, int n_in, const int *link_inter,
            int n_inter, double B, double _Complex *H) {
    int ns = ko_ns(lat);
    int ncell = lat->n1 * lat->n2 / 2;
    double _Complex *T = (double _Complex *)calloc((size_t)ns * ns, sizeof(double _Complex));
    for (int c = 1; c <= ncell; c++) {                    /* :254-268 in-cell links */
        for (int s1 = (c - 1) * 6 + 1; s1 <= c * 6; s1++)
            for (int s2 = (c - 1) * 6 + 1; s2 <= c * 6; s2++) {
                if (s1 >= s2) continue;
                int l1 = (s1 - 1) % 6 + 1, l2 = (s2 - 1) % 6 + 1, val;
                if (find_link_in(link_in, n_in, l1, l2, &val)) {
                    double r1[2], r2[2];
                    ko_get_site_coord(lat, s1, r1);
                    ko_get_site_coord(lat, s2, r2);
                    double phase = (B / 2.0) * (r1[0] + r2[0]) * (r2[1] - r1[1]);
                    T[(size_t)(s2 - 1) * ns + (s1 - 1)] = (double)val * cexp(I * phase);
                }
            }
    }
    for (int c1 = 1; c1 <= ncell; c1++)                   /* :270-280 inter-cell links */
        for (int c2 = 1; c2 <= ncell; c2++) {
            if (c1 == c2) continue;
            for (int s1 = (c1 - 1) * 6 + 1; s1 <= c1 * 6; s1++)
                for (int s2 = (c2 - 1) * 6 + 1; s2 <= c2 * 6; s2++) {
                    if (s1 >= s2) continue;
                    ko_apply_boundary_conditions(T, ns, lat, s1, s2, link_inter, n_inter, B);
                }
        }
    for (int i = 0; i < ns; i++)                          /* :281-287 */
        for (int j = 0; j < i; j++)
            if (T[(size_t)j * ns + i] != 0.0) { free(T); return -3; }
    for (int j = 0; j < ns; j++)                          /* :288  -(T + T') */
        for (int i = 0; i < ns; i++)
            H[(size_t)j * ns + i] = -(T[(size_t)j * ns + i] + conj(T[(size_t)i * ns + j]));
    free(T);
    return 0;
}

/* src/Hamiltonian.jl:447-451 get_nn: findall(!iszero, UpperTriangular(H)) in
 * column-major order -> (i, j) 1-based pairs with i <= j.  returns n_bonds. */
int ko_get_nn(const double _Complex *H, int ns, int32_t *bonds, int max_bonds) {
    int n = 0;
    for (int j = 0; j < ns; j++)
        for (int i = 0; i <= j; i++)
            if (H[(size_t)j * ns + i] != 0.0) {
                if (n < max_bonds) { bonds[2 * n] = i + 1; bonds[2 * n + 1] = j + 1; }
                n++;
            }
    return n;
}

/* ------------------------------------------------------------------ */
/* RNG: Julia Random.Xoshiro (xoshiro256++), SURVEY Appendix A.2       */
/* (stdlib, not under /root/reference: parity unpinned)                */
/* ------------------------------------------------------------------ */
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

uint64_t ko_xoshiro_next(uint64_t s[4]) {
    uint64_t res = rotl64(s[0] + s[3], 23) + s[0];
    uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl64(s[3], 45);
    return res;
}
/* rand(rng)::Float64 */
double ko_rand_f64(uint64_t s[4]) { return (double)(ko_xoshiro_next(s) >> 11) * 0x1.0p-53; }
/* rand(rng, 1:n): Julia SamplerRangeNDL (nearly division-less); 1-based result.
 * One draw is consumed even for n == 1.  (StatsBase.sample(rng, a) = a[rand(rng, 1:length(a))]) */
int64_t ko_rand_index(uint64_t s[4], uint64_t n) {
    uint64_t x = ko_xoshiro_next(s);
    unsigned __int128 m = (unsigned __int128)x * n;
    uint64_t l = (uint64_t)m;
    if (l < n) {
        uint64_t t = (0 - n) % n;
        while (l < t) {
            x = ko_xoshiro_next(s);
            m = (unsigned __int128)x * n;
            l = (uint64_t)m;
        }
    }
    return (int64_t)(m >> 64) + 1;
}
/* SplitMix64: used by tests/bench to derive per-walker Xoshiro states from a seed
 * (the host is the owner of the seeding policy; Julia's SHA-based seeding is not restated). */
uint64_t ko_splitmix64(uint64_t *x) {
    uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* ------------------------------------------------------------------ */
/* integer parts of the MC path (scalar-type independent)              */
/* ------------------------------------------------------------------ */
/* src/MonteCarlo.jl:460-474 Z.  bonds = 1-based (site1, site2) pairs */
int ko_Z(const int32_t *bonds, int n_bonds, const int64_t *kup, const int64_t *kdn) {
    int count = 0;
    for (int b = 0; b < n_bonds; b++) {
        int s1 = bonds[2 * b] - 1, s2 = bonds[2 * b + 1] - 1;
        if (kup[s1] != 0 && kdn[s2] != 0) count += 1;
        else if (kup[s2] != 0 && kdn[s1] != 0) count += 1;
    }
    return count;
}

/* src/Hamiltonian.jl:531-565 Sz.  returns 0 and *out = +-0.5, or -1 (doubly occupied),
 * -2 (unoccupied) [ArgumentError], -3 BoundsError */
int ko_Sz(int i, const int64_t *kup, const int64_t *kdn, int n, double *out) {
    if (!(1 <= i && i <= n)) return -3;
    int up = kup[i - 1] != 0, dn = kdn[i - 1] != 0;
    if (up && !dn) { *out = 0.5; return 0; }
    if (!up && dn) { *out = -0.5; return 0; }
    if (up && dn) return -1;
    return -2;
}

/* src/Hamiltonian.jl:636-672 spinInteraction! + :593-604 SzInteraction! + :711-720 getxprime,
 * flattened: for each bond emits up to two flip keys (K_up, l_up, K_down, l_down) with
 * coefficient -1/2, and accumulates the diagonal Sz_i*Sz_j sum.  keys: [max_keys][4] 1-based.
 * Distinct bonds always give distinct keys (the key contains both sites), so this flat list is
 * the reference's Dict.  returns number of keys or <0 (Sz error code). */
int ko_getxprime(const int32_t *bonds, int n_bonds, const int64_t *kup, const int64_t *kdn,
                 int ns, int64_t *keys, double *coefs, int max_keys, double *diag) {
    int nk = 0;
    double d = 0.0;
    for (int b = 0; b < n_bonds; b++) {
        int i = bonds[2 * b], j = bonds[2 * b + 1];
        int64_t i_up = kup[i - 1], j_up = kup[j - 1], i_dn = kdn[i - 1], j_dn = kdn[j - 1];
        if (j_up != 0 && i_dn != 0) {                     /* :649-657 */
            if (nk < max_keys) {
                keys[4 * nk] = i; keys[4 * nk + 1] = j_up; keys[4 * nk + 2] = j; keys[4 * nk + 3] = i_dn;
                coefs[nk] = -1.0 / 2.0;
            }
            nk++;
        }
        if (i_up != 0 && j_dn != 0) {                     /* :661-669 */
            if (nk < max_keys) {
                keys[4 * nk] = j; keys[4 * nk + 1] = i_up; keys[4 * nk + 2] = i; keys[4 * nk + 3] = j_dn;
                coefs[nk] = -1.0 / 2.0;
            }
            nk++;
        }
        double si, sj;
        int e;
        if ((e = ko_Sz(i, kup, kdn, ns, &si)) != 0) return e;
        if ((e = ko_Sz(j, kup, kdn, ns, &sj)) != 0) return e;
        d += si * sj;                                     /* :508-510 */
    }
    *diag = d;
    return nk;
}

/* src/MonteCarlo.jl:497-511 update_configurations! (integer part only) */
void ko_update_kappa(int64_t *kup, int64_t *kdn, int flag, int i, int site, int64_t l_up,
                     int64_t l_dn) {
    if (flag == 1) {
        kup[i - 1] = 0; kup[site - 1] = l_up;
        kdn[i - 1] = l_dn; kdn[site - 1] = 0;
    } else {
        kup[i - 1] = l_up; kup[site - 1] = 0;
        kdn[i - 1] = 0; kdn[site - 1] = l_dn;
    }
}

/* ------------------------------------------------------------------ */
/* scalar-typed MC core, instantiated twice                            */
/* ------------------------------------------------------------------ */
#define SCALAR double _Complex
#define SFX(name) name##_c128
#define S_ABS2(x) (creal(x) * creal(x) + cimag(x) * cimag(x))
#define S_ABS1(x) (fabs(creal(x)) + fabs(cimag(x)))
#define S_REAL(x) creal(x)
#define S_ISFINITE(x) (isfinite(creal(x)) && isfinite(cimag(x)))
#define S_FMA(a, b, c) ((a) * (b) + (c))
#include "oracle_mc_core.inc"
#undef SCALAR
#undef SFX
#undef S_ABS2
#undef S_ABS1
#undef S_REAL
#undef S_ISFINITE
#undef S_FMA

#define SCALAR double
#define SFX(name) name##_f64
#define S_ABS2(x) ((x) * (x))
#define S_ABS1(x) fabs(x)
#define S_REAL(x) (x)
#define S_ISFINITE(x) isfinite(x)
#define S_FMA(a, b, c) fma((a), (b), (c))
#include "oracle_mc_core.inc"
