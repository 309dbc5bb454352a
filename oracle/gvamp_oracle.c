/* oracle/gvamp_oracle.c -- TEST INFRASTRUCTURE (the parity oracle), never part of the product path.
 *
 * Plain-C restatement of the gVAMP reference's bed hot path, written from the reference's
 * *behaviour* (file:line cited per function, paths relative to /root/reference).  It is validated
 * against the unmodified reference compiled into oracle/_ref (tests/test_oracle_vs_ref.py, run in
 * the build container) and against the committed golden vectors in tests/golden/.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 *
 * Conventions: SNP-major PLINK .bed payload (no 3-byte magic): marker j occupies
 * bed[j*mbytes .. (j+1)*mbytes), 4 individuals per byte, least-significant pair first.
 * The reference's lookup tables (dotp_lut.hpp, na_lut.hpp) are replaced by their closed form
 * (SURVEY.md section 2.3): code 00 -> a=2,b=1 ; 01 -> a=0,b=0 (missing) ; 10 -> a=1,b=1 ; 11 -> a=0,b=1.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline double lut_a(unsigned c) { return c == 0u ? 2.0 : (c == 2u ? 1.0 : 0.0); } /* dotp_lut.hpp:3 */
static inline double lut_b(unsigned c) { return c == 1u ? 0.0 : 1.0; }                 /* dotp_lut.hpp:1030 */
static inline double lut_na(unsigned mask_nibble, int k) { return (double)((mask_nibble >> k) & 1u); } /* na_lut.hpp:3 */

/* Per-marker genotype-code counts.  counts[j*8 + c] = number of individuals i < N with code c and
 * phenotype present (mask bit set); counts[j*8 + 4 + c] = the same without the phenotype mask.
 * These integers are the bit-exact targets for the CUDA stats kernel (BASELINE north_star:
 * "Genotype decoding and NA counts must be bit-exact"). */
void orc_counts(const uint8_t* bed, long M, long mbytes, long N, const uint8_t* mask4, int64_t* counts) {
#pragma omp parallel for schedule(static)
    for (long j = 0; j < M; j++) {
        const uint8_t* col = bed + (size_t)j * (size_t)mbytes;
        int64_t c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (long p = 0; p < mbytes; p++) {
            for (int k = 0; k < 4; k++) {
                long i = 4 * p + k;
                if (i >= N) break;
                unsigned code = (col[p] >> (2 * k)) & 3u;
                c[4 + code]++;
                if ((mask4[p] >> k) & 1u) c[code]++;
            }
        }
        memcpy(counts + 8 * j, c, sizeof(c));
    }
}

/* data::compute_markers_statistics, scalar (non-MANVECT) branch: data.cpp:447-485.
 * mave = sum(a*na)/sum(b*na) (0 if no observation); msig = 1/sqrt(SS/(nonas-1)) [^alpha_scale],
 * 1 if SS == 0.  Loops run over all im4 = mbytes bytes including pad individuals, which only the
 * mask removes -- exactly like the reference. */
void orc_stats(const uint8_t* bed, long M, long mbytes, const uint8_t* mask4, int nonas, double alpha_scale,
               double* mave, double* msig) {
#pragma omp parallel for schedule(static)
    for (long j = 0; j < M; j++) {
        const uint8_t* col = bed + (size_t)j * (size_t)mbytes;
        double suma = 0.0, sumb = 0.0;
        for (long p = 0; p < mbytes; p++)
            for (int k = 0; k < 4; k++) {
                unsigned code = (col[p] >> (2 * k)) & 3u;
                suma += lut_a(code) * lut_na(mask4[p], k);
                sumb += lut_b(code) * lut_na(mask4[p], k);
            }
        double ave = (sumb != 0.0) ? suma / sumb : 0.0;
        mave[j] = ave;
        double sumsqr = 0.0;
        for (long p = 0; p < mbytes; p++)
            for (int k = 0; k < 4; k++) {
                unsigned code = (col[p] >> (2 * k)) & 3u;
                double val = (lut_a(code) - ave) * lut_b(code) * lut_na(mask4[p], k);
                sumsqr += val * val;
            }
        if (sumsqr != 0.0) {
            if (alpha_scale == 1.0)
                msig[j] = 1.0 / sqrt(sumsqr / ((double)nonas - 1.0));
            else
                msig[j] = 1.0 / pow(sqrt(sumsqr / ((double)nonas - 1.0)), alpha_scale);
        } else {
            msig[j] = 1.0;
        }
    }
}

/* data::dot_product (scalar branch data.cpp:758-779) + data::ATx (data.cpp:814-835).
 * u has 4*LB entries and is indexed relative to byte SB; no phenotype mask is applied inside. */
void orc_ATx(const uint8_t* bed, long M, long mbytes, long N, const double* u, const double* mave, const double* msig,
             long SB, long LB, double* out) {
    double scale = (SB == 0 && LB == mbytes) ? 1.0 / sqrt((double)N) : 1.0 / sqrt((double)(4 * LB));
#pragma omp parallel for schedule(static)
    for (long j = 0; j < M; j++) {
        const uint8_t* col = bed + (size_t)j * (size_t)mbytes;
        double dpa = 0.0, dpb = 0.0;
        for (long p = SB; p < SB + LB; p++)
            for (int k = 0; k < 4; k++) {
                unsigned code = (col[p] >> (2 * k)) & 3u;
                dpa += lut_a(code) * u[(p - SB) * 4 + k];
                dpb += lut_b(code) * u[(p - SB) * 4 + k];
            }
        out[j] = msig[j] * (dpa - mave[j] * dpb) * scale;
    }
}

/* data::Ax, scalar branch data.cpp:944-1007 (mask applied, single rank so the MPI_Allreduce at :995
 * is the identity).  out has 4*LB entries.  Markers are accumulated in index order like the
 * reference; the parallel loop is over bytes so every out[] element sees the same order of adds. */
void orc_Ax(const uint8_t* bed, long M, long mbytes, long N, const double* v, const double* mave, const double* msig,
            const uint8_t* mask4, long SB, long LB, double* out) {
    double scale = (SB == 0 && LB == mbytes) ? 1.0 / sqrt((double)N) : 1.0 / sqrt((double)(4 * LB));
#pragma omp parallel for schedule(static)
    for (long p = SB; p < SB + LB; p++) {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (long j = 0; j < M; j++) {
            unsigned byte = bed[(size_t)j * (size_t)mbytes + (size_t)p];
            double ave = mave[j];
            double sig_phen = msig[j] * v[j];
            for (int k = 0; k < 4; k++) {
                unsigned code = (byte >> (2 * k)) & 3u;
                acc[k] += (lut_a(code) - ave) * sig_phen * lut_b(code) * lut_na(mask4[p], k);
            }
        }
        for (int k = 0; k < 4; k++) out[(p - SB) * 4 + k] = acc[k] * scale;
    }
}

/* data::filter_pheno data.cpp:1065-1079: y with phenotype-NA entries set to 0. */
void orc_filter_pheno(const double* phen, long N, const uint8_t* mask4, double* out) {
    for (long i = 0; i < N; i++) out[i] = ((mask4[i / 4] >> (i % 4)) & 1u) ? phen[i] : 0.0;
}

/* ------------------------------------------------------------------------------------------------
 * Synthetic genotype generator: g_ij ~ Binomial(2, p_j), p_j ~ U(0.01, 0.5),
 * PLINK encoding 2 -> 00, 1 -> 10, 0 -> 11, missing -> 01, pad individuals -> 00.  The stream is a
 * counter-based hash keyed by (seed, global marker index, individual block) so that any marker slice
 * can be regenerated independently; gvamp_b200/csrc/synth.cu implements the identical function on
 * the device and tests compare the two byte for byte.
 * ------------------------------------------------------------------------------------------------ */
static inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
uint64_t orc_marker_key(uint64_t seed, uint64_t j) { return mix64(mix64(seed) ^ ((j + 1) * 0xD1B54A32D192ED03ull)); }
double orc_marker_freq(uint64_t seed, uint64_t j) {
    uint64_t hj = orc_marker_key(seed, j);
    return 0.01 + 0.49 * ((double)(hj >> 11) * (1.0 / 9007199254740992.0));
}
/* bytes for markers [j0, j0+M) of a population of N individuals; miss_thr16 = round(rate * 65536) */
void orc_synth_bed(uint64_t seed, long j0, long M, long N, unsigned miss_thr16, uint8_t* out) {
    long mbytes = (N + 3) / 4;
#pragma omp parallel for schedule(static)
    for (long jj = 0; jj < M; jj++) {
        uint64_t hj = orc_marker_key(seed, (uint64_t)(j0 + jj));
        double pj = 0.01 + 0.49 * ((double)(hj >> 11) * (1.0 / 9007199254740992.0));
        unsigned thr = (unsigned)(pj * 65536.0);
        uint8_t* col = out + (size_t)jj * (size_t)mbytes;
        for (long p = 0; p < mbytes; p++) {
            uint64_t h0 = mix64(hj + 0x9E3779B97F4A7C15ull * (uint64_t)(2 * p + 1));
            uint64_t h1 = mix64(hj + 0x9E3779B97F4A7C15ull * (uint64_t)(2 * p + 2));
            uint64_t hm = miss_thr16 ? mix64((hj ^ 0xA5A5A5A5A5A5A5A5ull) + 0xC2B2AE3D27D4EB4Full * (uint64_t)(p + 1)) : 0;
            unsigned byte = 0;
            for (int k = 0; k < 4; k++) {
                long i = 4 * p + k;
                unsigned code;
                if (i >= N) {
                    code = 0u;
                } else {
                    uint64_t h = (k < 2) ? h0 : h1;
                    int sh = (k & 1) * 32;
                    unsigned d = (((unsigned)(h >> sh) & 0xFFFFu) < thr) + (((unsigned)(h >> (sh + 16)) & 0xFFFFu) < thr);
                    code = d == 2u ? 0u : (d == 1u ? 2u : 3u);
                    if (miss_thr16 && (((unsigned)(hm >> (16 * k)) & 0xFFFFu) < miss_thr16)) code = 1u;
                }
                byte |= code << (2 * k);
            }
            col[p] = (uint8_t)byte;
        }
    }
}
