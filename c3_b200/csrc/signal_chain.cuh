// On-device signal generation chain: from pulse parameters straight to the
// control fields signals[B,K,N] the propagator kernels read, for a whole batch of parameter samples.
//
// One kernel replaces, per (sample, drive line), the reference's device chain
//   LO (c3/generator/devices.py:1063-1130, noise-free branch)
//   AWG.create_IQ (:1159-1197) -> Instruction.get_awg_signal (c3/signal/gates.py:341-370)
//        -> Envelope shape values with mask / t_before offset / DRAG quadrature (c3/signal/pulse.py:93-180,
//           c3/libraries/envelopes.py: no_drive, rect, flattop, gaussian_sigma, cosine, gaussian_nonorm)
//   DigitalToAnalog (:296-351, nearest-neighbour resampling with half-pixel centres)
//   Response / ResponseFFT (:585-701; the FFT products of c3/utils/tf_utils.py:441-515 are a 30-tap direct
//        convolution here: out[n] = sum_m r[m] x[n - 1 - m] (legacy) or x[n - m])
//   Mixer (:906-939) and VoltsToHertz (:187-221) or FluxTuning (:480-529)
// so no intermediate [N] array ever leaves the SM and the 8 K N bytes per sample are written once.
#pragma once
#include "c3b_common.cuh"

namespace c3b {

// layout of one envelope's parameter row and of one drive line's chain row (doubles)
enum { ENV_AMP = 0, ENV_TFINAL, ENV_SIGMA, ENV_XY, ENV_FREQ_OFFSET, ENV_DELTA, ENV_TUP, ENV_TDOWN, ENV_RISEFALL, ENV_NPAR };
enum { CH_SIM_RES = 0, CH_AWG_RES, CH_RISE_TIME, CH_RESP_KIND, CH_OUT_KIND, CH_V2HZ, CH_PHI, CH_PHI0, CH_OMEGA0, CH_ANHAR, CH_D, CH_NPAR };
enum { SHAPE_NO_DRIVE = 0, SHAPE_RECT, SHAPE_GAUSSIAN_NONORM, SHAPE_GAUSSIAN_SIGMA, SHAPE_COSINE, SHAPE_FLATTOP,
       SHAPE_TRAPEZOID, SHAPE_FLATTOP_RISEFALL, SHAPE_GAUSSIAN_DER_NONORM, SHAPE_GAUSSIAN_DER, SHAPE_DRAG_SIGMA, SHAPE_DRAG_DER,
       // "extended" shapes: grid-dependent (normalised by their maximum over the AWG grid, defined by sample index) or
       // parametrised by arrays (table row of the envelope, SignalParams::table); forward only
       SHAPE_FIRST_EXT,
       SHAPE_FLATTOP_CUT = SHAPE_FIRST_EXT,   // envelopes.py:281-302   erf product clipped to [0,1], over its grid maximum
       SHAPE_FLATTOP_CUT_CENTER,              // :305-327   width in the sigma slot; clipped to [0,2], not normalised
       SHAPE_FLATTOP_VARIANT,                 // :565-587   ramp in the sigma slot
       SHAPE_COSINE_FLATTOP,                  // :440-466   t_rise in the sigma slot; defined on sample indices
       SHAPE_DELTA_PULSE,                     // :128-139   table: M, t_sig[M]
       SHAPE_PWC,                             // :31-34     table: M, inphase[M], quadrature[M]  (complex, by sample index)
       SHAPE_PWC_SHAPE,                       // :37-68     table: M, t_bin_start, t_bin_end, inphase[M]  (linear interpolation)
       SHAPE_PWC_SYMMETRIC,                   // :104-125   same, mirrored at t_final / 2
       SHAPE_PWC_SHAPE_PLATEAU,               // :71-101    same + width (< 0: none)
       SHAPE_FOURIER_SIN,                     // :142-168   table: M, amps[M], freqs[M], phases[M]
       SHAPE_FOURIER_COS,                     // :171-191   table: M, amps[M], freqs[M]
       SHAPE_SLEPIAN_FOURIER,                 // :330-363   table: width, offset, risefall (< 0: none), M, coeffs[M], S, sin_coeffs[S]
       SHAPE_END };

struct SignalParams {
    const double* env;       // [B, K, E, ENV_NPAR]
    const int* shape;        // [K, E]  shape id, < 0: unused slot
    const int* flags;        // [K, E]  bit 0: DRAG quadrature, bit 1: subtract the value one sample before the start
    const double* lo_freq;   // [B, K]  carrier angular frequency
    const double* chain;     // [K, CH_NPAR] or [B, K, CH_NPAR]
    int chain_batched;
    double t_start, t_end;
    int B, K, E, N, max_awg, max_taps;
    double* out;             // [B, K, N]
    // noise devices (c3/generator/devices.py:943-1035), all optional (noise == nullptr: noise-free chain)
    const double* noise;     // [K, NOISE_NPAR] or [B, K, NOISE_NPAR]: see the NOISE_* enum
    int noise_batched;
    unsigned long long seed; // one noise realisation per (seed, batch row, line): counter-based, reproducible
    double* noise_out;       // [B, K, NOISE_NTRACE, N] realised noise traces (what Device.signal["noise"] holds) or null
    const double* table;     // [K, E, T] array parameters of the extended shapes (shared by the batch) or null
    int T;
};

// one drive line's noise row and the layout of the realised traces
enum { NOISE_AWG_AMP = 0,    // Additive_Noise after the AWG: amp * N(0,1) on every in-phase and quadrature AWG sample
       NOISE_LO_PERC,        // LONoise: perc * N(0,1) on the LO's cos and sin at every simulation sample
       NOISE_ADD_AMP,        // Additive_Noise after the mixer: amp * N(0,1) per simulation sample
       NOISE_DC_AMP,         // DC_Noise: ONE amp * N(0,1) offset per realisation
       NOISE_PINK_AMP,       // Pink_Noise: amp * (sum of bfl_num bistable fluctuators)
       NOISE_PINK_BFL,       //             number of fluctuators (<= 32)
       NOISE_DC_OFFSET,      // DC_Offset: deterministic offset added after the mixer
       NOISE_NPAR };
enum { TRACE_AWG_I = 0, TRACE_AWG_Q, TRACE_LO_COS, TRACE_LO_SIN, TRACE_ADD, TRACE_DC, TRACE_PINK, NOISE_NTRACE };
enum { STREAM_AWG = 1, STREAM_LO, STREAM_ADD, STREAM_DC, STREAM_PINK_INIT, STREAM_PINK_FLIP };

// ---- Philox4x32-10 (Salmon et al., SC'11): counter-based, so every (realisation, line, stream, sample) has its own
// random numbers without any state or ordering between threads; restated in oracle/c3_noise_oracle.py
__device__ __forceinline__ void philox4x32_10(unsigned int k0, unsigned int k1, unsigned int c0, unsigned int c1, unsigned int c2,
                                              unsigned int c3, unsigned int (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned int hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned int hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const unsigned int n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// random words of (seed, line bk, stream, index)
__device__ __forceinline__ void noise_words(unsigned long long seed, int bk, int stream, unsigned int idx, unsigned int (&w)[4]) {
    philox4x32_10((unsigned int)seed, (unsigned int)(seed >> 32), idx, (unsigned int)bk, (unsigned int)stream, 0u, w);
}
__device__ __forceinline__ double u01(unsigned int hi, unsigned int lo) {      // 53-bit uniform in (0, 1)
    const unsigned long long m = ((unsigned long long)(hi >> 5) << 26) | (lo >> 6);
    return ((double)m + 0.5) * (1.0 / 9007199254740992.0);
}
// two independent standard normals (Box-Muller)
__device__ __forceinline__ void noise_normals(unsigned long long seed, int bk, int stream, unsigned int idx, double& z0, double& z1) {
    unsigned int w[4];
    noise_words(seed, bk, stream, idx, w);
    const double r = sqrt(-2.0 * log(u01(w[0], w[1])));
    double sn, cs;
    sincos(2.0 * M_PI * u01(w[2], w[3]), &sn, &cs);
    z0 = r * cs;
    z1 = r * sn;
}

// ---- forward-mode dual numbers: the SAME envelope formulas give values (T = double) and parameter derivatives
// (T = Dual, seeded on one of the 9 envelope parameters) -- the reference differentiates them with tf.GradientTape
struct Dual {
    double v, d;
    __device__ __forceinline__ Dual() : v(0.0), d(0.0) {}
    __device__ __forceinline__ Dual(double v_) : v(v_), d(0.0) {}
    __device__ __forceinline__ Dual(double v_, double d_) : v(v_), d(d_) {}
};
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
__device__ __forceinline__ Dual operator-(Dual a) { return Dual(-a.v, -a.d); }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, a.d * b.v + a.v * b.d); }
__device__ __forceinline__ Dual operator/(Dual a, Dual b) { return Dual(a.v / b.v, (a.d * b.v - a.v * b.d) / (b.v * b.v)); }
__device__ __forceinline__ double t_val(double x) { return x; }
__device__ __forceinline__ double t_val(Dual x) { return x.v; }
__device__ __forceinline__ double t_exp(double x) { return exp(x); }
__device__ __forceinline__ Dual t_exp(Dual x) { const double e = exp(x.v); return Dual(e, e * x.d); }
__device__ __forceinline__ double t_erf(double x) { return erf(x); }
__device__ __forceinline__ Dual t_erf(Dual x) { return Dual(erf(x.v), 1.1283791670955126 * exp(-x.v * x.v) * x.d); }
__device__ __forceinline__ double t_cos(double x) { return cos(x); }
__device__ __forceinline__ Dual t_cos(Dual x) { return Dual(cos(x.v), -sin(x.v) * x.d); }
__device__ __forceinline__ double t_sin(double x) { return sin(x); }
__device__ __forceinline__ Dual t_sin(Dual x) { return Dual(sin(x.v), cos(x.v) * x.d); }
__device__ __forceinline__ double t_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ Dual t_sqrt(Dual x) { const double r = sqrt(x.v); return Dual(r, 0.5 * x.d / r); }
__device__ __forceinline__ double t_sigmoid(double x) { return 1.0 / (1.0 + exp(-x)); }
__device__ __forceinline__ Dual t_sigmoid(Dual x) { const double s = 1.0 / (1.0 + exp(-x.v)); return Dual(s, s * (1.0 - s) * x.d); }

template <typename T>
__device__ __forceinline__ T shape_value(int id, T t, const T* e) {
    switch (id) {
        case SHAPE_RECT: return T(1.0);
        case SHAPE_GAUSSIAN_NONORM: {
            const T u = t - e[ENV_TFINAL] / T(2.0), s = e[ENV_SIGMA];
            return t_exp(-(u * u) / (T(2.0) * s * s));
        }
        case SHAPE_GAUSSIAN_SIGMA: {
            const T tf = e[ENV_TFINAL], s = e[ENV_SIGMA], u = t - tf / T(2.0);
            const T gauss = t_exp(-(u * u) / (T(2.0) * s * s));
            const T offset = t_exp(-(tf * tf) / (T(8.0) * s * s));
            const T norm = t_sqrt(T(2 * M_PI) * s * s) * t_erf(tf / (T(sqrt(8.0)) * s)) - tf * offset;
            return (gauss - offset) / norm;
        }
        case SHAPE_COSINE: return T(0.5) * (T(1.0) - t_cos(T(2 * M_PI) * t / e[ENV_TFINAL]));
        case SHAPE_FLATTOP:
            return (T(1.0) + t_erf((t - e[ENV_TUP]) / e[ENV_RISEFALL])) / T(2.0) *
                   (T(1.0) + t_erf((-t + e[ENV_TDOWN]) / e[ENV_RISEFALL])) / T(2.0);
        case SHAPE_TRAPEZOID: {             // envelopes.py:200-224: linear slopes of width 2.5 risefall (the fall wins where both apply)
            const T w = e[ENV_RISEFALL] * T(2.5), tf = e[ENV_TFINAL];
            if (t_val(t) >= t_val(tf - w)) return (tf - t) / w;
            if (t_val(t) <= t_val(w)) return t / w;
            return T(1.0);
        }
        case SHAPE_FLATTOP_RISEFALL: {      // envelopes.py:227-250: flattop with t_up = risefall, t_down = t_final - risefall
            const T rf = e[ENV_RISEFALL];
            return (T(1.0) + t_erf((t - rf) / rf)) / T(2.0) * (T(1.0) + t_erf((-t + e[ENV_TFINAL] - rf) / rf)) / T(2.0);
        }
        case SHAPE_GAUSSIAN_DER_NONORM:     // envelopes.py:490-500
        case SHAPE_GAUSSIAN_DER: {          // :503-516 (same, over the norm of gaussian_sigma)
            const T tf = e[ENV_TFINAL], s = e[ENV_SIGMA], u = t - tf / T(2.0);
            T g = t_exp(-(u * u) / (T(2.0) * s * s)) * u / (s * s);
            if (id == SHAPE_GAUSSIAN_DER)       // sqrt(8) in float32 as the reference has it here (envelopes.py:514)
                g = g / (t_sqrt(T(2 * M_PI) * s * s) * t_erf(tf / (T(2.8284270763397217) * s)) - tf * t_exp(-(tf * tf) / (T(8.0) * s * s)));
            return g;
        }
        case SHAPE_DRAG_SIGMA:              // envelopes.py:519-530: (gauss - offset)^2 / norm
        case SHAPE_DRAG_DER: {              // :545-562: -2 (gauss - offset) gauss (t - t_final/2) / sigma^2 / norm
            const T tf = e[ENV_TFINAL], s = e[ENV_SIGMA], u = t - tf / T(2.0);
            const T gauss = t_exp(-(u * u) / (T(2.0) * s * s));
            const T offset = t_exp(-(tf * tf) / (T(8.0) * s * s));
            const T norm = t_sqrt(T(2 * M_PI) * s * s) * t_erf(tf / (T(sqrt(8.0)) * s)) - tf * offset;
            if (id == SHAPE_DRAG_SIGMA) return (gauss - offset) * (gauss - offset) / norm;
            return T(-2.0) * (gauss - offset) * gauss * u / (s * s) / norm;
        }
        default: return T(0.0);
    }
}

template <typename T>
__device__ __forceinline__ T shape_deriv(int id, T t, const T* e) {
    switch (id) {
        case SHAPE_GAUSSIAN_NONORM:
        case SHAPE_GAUSSIAN_SIGMA: {
            const T tf = e[ENV_TFINAL], s = e[ENV_SIGMA], u = t - tf / T(2.0);
            T g = t_exp(-(u * u) / (T(2.0) * s * s)) * (-u / (s * s));
            if (id == SHAPE_GAUSSIAN_SIGMA) {
                const T offset = t_exp(-(tf * tf) / (T(8.0) * s * s));
                g = g / (t_sqrt(T(2 * M_PI) * s * s) * t_erf(tf / (T(sqrt(8.0)) * s)) - tf * offset);
            }
            return g;
        }
        case SHAPE_COSINE: return T(0.5) * t_sin(T(2 * M_PI) * t / e[ENV_TFINAL]) * T(2 * M_PI) / e[ENV_TFINAL];
        case SHAPE_FLATTOP:
        case SHAPE_FLATTOP_RISEFALL: {
            const T rf = e[ENV_RISEFALL];
            const T tu = (id == SHAPE_FLATTOP) ? e[ENV_TUP] : rf, td = (id == SHAPE_FLATTOP) ? e[ENV_TDOWN] : e[ENV_TFINAL] - rf;
            const T up = (t - tu) / rf, dn = (-t + td) / rf;
            const T c = T(2 / sqrt(M_PI)) / rf;
            return (c * t_exp(-(up * up)) * (T(1.0) + t_erf(dn)) - (T(1.0) + t_erf(up)) * c * t_exp(-(dn * dn))) / T(4.0);
        }
        case SHAPE_TRAPEZOID: {
            const T w = e[ENV_RISEFALL] * T(2.5), tf = e[ENV_TFINAL];
            if (t_val(t) >= t_val(tf - w)) return T(-1.0) / w;
            if (t_val(t) <= t_val(w)) return T(1.0) / w;
            return T(0.0);
        }
        case SHAPE_GAUSSIAN_DER_NONORM:
        case SHAPE_GAUSSIAN_DER: {          // d/dt [ gauss u / s^2 ] = gauss (1 / s^2 - u^2 / s^4)
            const T tf = e[ENV_TFINAL], s = e[ENV_SIGMA], u = t - tf / T(2.0), s2 = s * s;
            T g = t_exp(-(u * u) / (T(2.0) * s2)) * (T(1.0) / s2 - u * u / (s2 * s2));
            if (id == SHAPE_GAUSSIAN_DER)
                g = g / (t_sqrt(T(2 * M_PI) * s2) * t_erf(tf / (T(2.8284270763397217) * s)) - tf * t_exp(-(tf * tf) / (T(8.0) * s2)));
            return g;
        }
        case SHAPE_DRAG_SIGMA:
        case SHAPE_DRAG_DER: {
            const T tf = e[ENV_TFINAL], s = e[ENV_SIGMA], u = t - tf / T(2.0), s2 = s * s;
            const T gauss = t_exp(-(u * u) / (T(2.0) * s2)), dg = -(u / s2) * gauss;
            const T offset = t_exp(-(tf * tf) / (T(8.0) * s2));
            const T norm = t_sqrt(T(2 * M_PI) * s2) * t_erf(tf / (T(sqrt(8.0)) * s)) - tf * offset;
            if (id == SHAPE_DRAG_SIGMA) return T(2.0) * (gauss - offset) * dg / norm;
            // d/dt [ (gauss - offset) gauss u ] = dg gauss u + (gauss - offset) dg u + (gauss - offset) gauss
            return T(-2.0) / (s2 * norm) * (dg * gauss * u + (gauss - offset) * dg * u + (gauss - offset) * gauss);
        }
        default: return T(0.0);
    }
}

// One envelope's contribution to the AWG sample at time t (offset from the instruction start):
// amp * (mask * shape [- value one sample before the start] - i * delta * mask * shape' * dt) * exp(i (xy - w_off t))
// (c3/signal/gates.py:341-370, c3/signal/pulse.py:93-180).  off0 / off1 are the first two grid offsets.
template <typename T>
__device__ __forceinline__ void awg_term(int id, int fl, double t, const T* ev, double off0, double off1, T& re, T& im) {
    const double dts = off1 - off0;
    const T tt(t);
    const T mask = t_sigmoid(T((t / dts + 0.001) * 1e6)) * t_sigmoid((T(0.999) * ev[ENV_TFINAL] - tt) / T(dts) * T(1e6));
    T sv = shape_value<T>(id, tt, ev);
    if (fl & 2) sv = sv - shape_value<T>(id, T(2 * off0 - off1), ev);
    const T env_re = mask * sv;
    const T env_im = (fl & 1) ? -(mask * shape_deriv<T>(id, tt, ev) * T(dts)) * ev[ENV_DELTA] : T(0.0);
    const T ph = ev[ENV_XY] - ev[ENV_FREQ_OFFSET] * tt;
    const T cs = t_cos(ph), sn = t_sin(ph);
    re = ev[ENV_AMP] * (env_re * cs - env_im * sn);
    im = ev[ENV_AMP] * (env_re * sn + env_im * cs);
}

// numpy.linspace(start, stop, num)[i]
__device__ __forceinline__ double linspace_at(double start, double stop, int num, int i) {
    if (num <= 1) return start;
    if (i == num - 1) return stop;
    return start + i * ((stop - start) / (num - 1));
}

__device__ __forceinline__ double flux_factor(double p, double phi0, double d, bool has_d) {
    const double x = M_PI * p / phi0;
    const double c = cos(x), s = sin(x);
    return has_d ? sqrt(sqrt(c * c + d * d * s * s)) : sqrt(fabs(c));
}

// per-(sample, line) view of the chain: grids, resampling ratio, response variant
struct ChainCtx {
    const double* ch;
    double sim_res, awg_res, rise, a0, a1, off0, off1, s0, s1, ratio, w_lo;
    int resp_kind, out_kind, n_awg, N, shift, taps;
    bool has_d;
};

__device__ __forceinline__ ChainCtx chain_ctx(const SignalParams& p, const int bk, const int k) {
    ChainCtx c;
    c.ch = p.chain + (p.chain_batched ? (size_t)bk : (size_t)k) * CH_NPAR;
    c.sim_res = c.ch[CH_SIM_RES]; c.awg_res = c.ch[CH_AWG_RES]; c.rise = c.ch[CH_RISE_TIME];
    c.resp_kind = (int)c.ch[CH_RESP_KIND]; c.out_kind = (int)c.ch[CH_OUT_KIND];
    const double span = fabs(p.t_start - p.t_end);
    c.n_awg = (int)(span * c.awg_res);
    c.N = p.N;
    const double dt_awg = 1.0 / c.awg_res, dt_sim = 1.0 / c.sim_res;
    c.a0 = p.t_start + dt_awg / 2; c.a1 = p.t_end - dt_awg / 2;
    c.off0 = linspace_at(c.a0, c.a1, c.n_awg, 0) - p.t_start;
    c.off1 = c.n_awg > 1 ? linspace_at(c.a0, c.a1, c.n_awg, 1) - p.t_start : c.off0 + dt_awg;
    c.s0 = p.t_start + dt_sim / 2; c.s1 = p.t_end - dt_sim / 2;
    c.ratio = (double)c.n_awg / (double)c.N;
    c.w_lo = p.lo_freq[bk];
    c.shift = (c.resp_kind == 1) ? 1 : 0;
    c.taps = (c.resp_kind != 0) ? (int)floor(c.rise * c.sim_res) : 0;
    c.has_d = c.ch[CH_D] == c.ch[CH_D];                  // NaN: symmetric SQUID
    return c;
}


// ---- extended shapes (forward only) ------------------------------------------------------------------------------------
// tfp.math.interp_regular_1d_grid(x, x_min, x_max, y[M], fill_value_below = fill_value_above = 0)
__device__ __forceinline__ double interp_regular(double x, double x_min, double x_max, const double* y, int M) {
    if (x < x_min || x > x_max) return 0.0;
    if (M == 1) return y[0];
    double u = (x - x_min) / (x_max - x_min) * (double)(M - 1);
    u = fmin(fmax(u, 0.0), (double)(M - 1));
    int lo = (int)floor(u);
    int hi = min(lo + 1, M - 1);
    lo = max(hi - 1, 0);
    const double w = u - (double)lo;
    return w * y[hi] + (1.0 - w) * y[lo];
}

__device__ __forceinline__ double slepian_x(double t, double t_final, double width, double risefall, double& length) {
    if (risefall >= 0.0) {
        const double plateau = width - risefall * 2;
        double x = t;
        if (t > (t_final + plateau) / 2) x = t - plateau / 2;
        if (t < (t_final - plateau) / 2) x = t + plateau / 2;
        if (fabs(t - t_final / 2) < plateau / 2) x = t_final / 2;
        length = risefall * 2;
        return x;
    }
    length = width;
    return t;
}

// value of an extended shape at AWG sample j (time offset t); im: the quadrature of the complex "pwc" shape.
// NOT normalised for FLATTOP_CUT / SLEPIAN_FOURIER (fill_awg divides by the grid maximum).
__device__ __forceinline__ double shape_value_ext(int id, double t, int j, const double* e, const double* tab, const ChainCtx& c,
                                                  double t_start, double& im) {
    im = 0.0;
    switch (id) {
        case SHAPE_FLATTOP_CUT: {
            const double v = erf((t - e[ENV_TUP]) / e[ENV_RISEFALL]) * erf((-t + e[ENV_TDOWN]) / e[ENV_RISEFALL]);
            return fmin(fmax(v, 0.0), 1.0);
        }
        case SHAPE_FLATTOP_CUT_CENTER: {
            const double t_up = e[ENV_TFINAL] / 2 - e[ENV_SIGMA] / 2, t_down = e[ENV_TFINAL] / 2 + e[ENV_SIGMA] / 2;
            const double v = erf((t - t_up) / e[ENV_RISEFALL]) * erf((-t + t_down) / e[ENV_RISEFALL]);
            return fmin(fmax(v, 0.0), 2.0);
        }
        case SHAPE_FLATTOP_VARIANT: {
            const double t_up = e[ENV_TUP], t_down = e[ENV_TDOWN];
            double ramp = e[ENV_SIGMA];
            if (ramp > (t_down - t_up) / 2) ramp = (t_down - t_up) / 2;
            const double sigma = sqrt(2.0) * ramp * 0.2;
            if (t_up <= t && t <= t_up + ramp) return exp(-((t - t_up - ramp) * (t - t_up - ramp)) / (2 * sigma * sigma));
            if (t_up + ramp < t && t < t_down - ramp) return 1.0;
            if (t_down >= t && t >= t_down - ramp) return exp(-((t - t_down + ramp) * (t - t_down + ramp)) / (2 * sigma * sigma));
            return 0.0;
        }
        case SHAPE_COSINE_FLATTOP: {           // rise over the first n_rise samples, the same n_rise time values again for the fall
            const double t_rise = e[ENV_SIGMA];
            const int n_rise = (int)(t_rise / (c.off1 - c.off0));
            const int n_flat = c.n_awg - 2 * n_rise;
            if (j < n_rise) return 0.5 * (1.0 - cos(M_PI * t / t_rise));
            if (j < n_rise + n_flat) return 1.0;
            const double tf = linspace_at(c.a0, c.a1, c.n_awg, j - n_rise - n_flat) - t_start;
            return 0.5 * (1.0 + cos(M_PI * tf / t_rise));
        }
        case SHAPE_DELTA_PULSE: {              // 1 on the grid point(s) closest to t_sig + 1 ns
            const int M = (int)tab[0];
            const double dts = c.off1 - c.off0;
            double v = 0.0;
            for (int m = 0; m < M; ++m) {
                const double ts = tab[1 + m];
                const double mine = (t - ts - 1e-9) * (t - ts - 1e-9);
                int j0 = (int)floor((ts + 1e-9 - c.off0) / dts);
                double best = 1e300;
                for (int q = j0 - 1; q <= j0 + 2; ++q) {
                    const int qq = min(max(q, 0), c.n_awg - 1);
                    const double tq = linspace_at(c.a0, c.a1, c.n_awg, qq) - t_start;
                    best = fmin(best, (tq - ts - 1e-9) * (tq - ts - 1e-9));
                }
                if (mine == best) v = 1.0;
            }
            return v;
        }
        case SHAPE_PWC: {
            const int M = (int)tab[0];
            if (j >= M) return 0.0;
            im = tab[1 + M + j];
            return tab[1 + j];
        }
        case SHAPE_PWC_SHAPE: return interp_regular(t, tab[1], tab[2], tab + 4, (int)tab[0]);
        case SHAPE_PWC_SYMMETRIC: return interp_regular(t > e[ENV_TFINAL] / 2 ? -t + e[ENV_TFINAL] : t, tab[1], tab[2], tab + 4, (int)tab[0]);
        case SHAPE_PWC_SHAPE_PLATEAU: {
            const double width = tab[3];
            if (width < 0.0) return interp_regular(t, tab[1], tab[2], tab + 4, (int)tab[0]);
            const double plateau = width - (tab[2] - tab[1]), t_mid = (tab[2] - tab[1]) / 2;
            double x = t;
            if (t > t_mid + plateau) x = t - plateau;
            if (t < t_mid) x = t;
            if (t < t_mid + plateau && t > t_mid) x = t_mid;
            if (x == t_mid) return 1.0;
            return interp_regular(x, tab[1], tab[2], tab + 4, (int)tab[0]);
        }
        case SHAPE_FOURIER_SIN: {
            const int M = (int)tab[0];
            double v = 0.0;
            for (int m = 0; m < M; ++m) v += tab[1 + m] * sin(tab[1 + M + m] * t + tab[1 + 2 * M + m]);
            return v;
        }
        case SHAPE_FOURIER_COS: {
            const int M = (int)tab[0];
            double v = 0.0;
            for (int m = 0; m < M; ++m) v += tab[1 + m] * cos(tab[1 + M + m] * t);
            return v;
        }
        case SHAPE_SLEPIAN_FOURIER: {
            const double t_final = e[ENV_TFINAL], width = tab[0];
            double length;
            const double x = slepian_x(t, t_final, width, tab[2], length);
            const int M = (int)tab[3];
            const int S = (int)tab[4 + M];
            double v = 0.0;
            for (int m = 0; m < M; ++m) v += tab[4 + m] * (1.0 - cos(2 * M_PI * (m + 1) * (x - (t_final - length) / 2) / length));
            for (int m = 0; m < S; ++m) v += tab[5 + M + m] * sin((M_PI * (2 * m + 1)) * (x - (t_final - length) / 2) / length);
            if (fabs(t_final / 2 - t) > width / 2) v = 0.0;
            return v;
        }
        default: return 0.0;
    }
}

__device__ __forceinline__ double block_max_128(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    return fmax(fmax(red[0], red[1]), fmax(red[2], red[3]));
}

// AWG samples: sum of the line's envelopes on the AWG time grid -> sI, sQ [n_awg].  Called by the whole CTA (barriers).
__device__ __forceinline__ void fill_awg(const SignalParams& p, const ChainCtx& c, const int b, const int k, double* sI, double* sQ) {
    __shared__ double s_red[4];
    for (int j = threadIdx.x; j < c.n_awg; j += blockDim.x) { sI[j] = 0.0; sQ[j] = 0.0; }
    for (int e = 0; e < p.E; ++e) {
        const int id = p.shape[k * p.E + e];                 // CTA-uniform
        if (id < 0) continue;
        const double* ev = p.env + (((size_t)b * p.K + k) * p.E + e) * ENV_NPAR;
        const int fl = p.flags[k * p.E + e];
        if (id < SHAPE_FIRST_EXT) {
            for (int j = threadIdx.x; j < c.n_awg; j += blockDim.x) {
                const double t = linspace_at(c.a0, c.a1, c.n_awg, j) - p.t_start;
                double tr, ti;
                awg_term<double>(id, fl, t, ev, c.off0, c.off1, tr, ti);
                sI[j] += tr;
                sQ[j] += ti;
            }
            continue;
        }
        const double* tab = p.table ? p.table + ((size_t)k * p.E + e) * p.T : nullptr;
        double norm = 1.0;
        if (id == SHAPE_FLATTOP_CUT || id == SHAPE_SLEPIAN_FOURIER) {       // shape /= reduce_max(shape) over the AWG grid
            double mx = -1e300, dummy;
            for (int j = threadIdx.x; j < c.n_awg; j += blockDim.x)
                mx = fmax(mx, shape_value_ext(id, linspace_at(c.a0, c.a1, c.n_awg, j) - p.t_start, j, ev, tab, c, p.t_start, dummy));
            norm = block_max_128(mx, s_red);
        }
        const double dts = c.off1 - c.off0;
        for (int j = threadIdx.x; j < c.n_awg; j += blockDim.x) {
            const double t = linspace_at(c.a0, c.a1, c.n_awg, j) - p.t_start;
            double qv;
            double sv = shape_value_ext(id, t, j, ev, tab, c, p.t_start, qv) / norm;
            if (id == SHAPE_SLEPIAN_FOURIER) sv = sv * (1.0 - tab[1] / ev[ENV_AMP]) + tab[1] / ev[ENV_AMP];
            const double mask = t_sigmoid((t / dts + 0.001) * 1e6) * t_sigmoid((0.999 * ev[ENV_TFINAL] - t) / dts * 1e6);
            const double env_re = mask * sv, env_im = mask * qv;
            const double ph = ev[ENV_XY] - ev[ENV_FREQ_OFFSET] * t;
            const double cs = cos(ph), sn = sin(ph);
            sI[j] += ev[ENV_AMP] * (env_re * cs - env_im * sn);
            sQ[j] += ev[ENV_AMP] * (env_re * sn + env_im * cs);
        }
    }
}

// normalised Gaussian rise function of the Response device -> sR [taps]; contains CTA barriers
__device__ __forceinline__ void fill_taps(const ChainCtx& c, double* sR, double* s_norm) {
    if (c.resp_kind != 0) {
        const double cen = (c.resp_kind == 2) ? (c.rise - 1 / c.sim_res) / 2 : (c.rise + 1 / c.sim_res) / 2;
        const double sg = c.rise / 4;
        const double offset = exp(-((-1 - cen) * (-1 - cen)) / (2 * sg * sg));
        for (int m = threadIdx.x; m < c.taps; m += blockDim.x) {
            const double tr = linspace_at(0.0, c.rise, c.taps, m);
            sR[m] = exp(-((tr - cen) * (tr - cen)) / (2 * sg * sg)) - offset;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int m = 0; m < c.taps; ++m) s += sR[m];
            *s_norm = s;
        }
        __syncthreads();
        for (int m = threadIdx.x; m < c.taps; m += blockDim.x) sR[m] = sR[m] / *s_norm;
    }
    __syncthreads();
}

__device__ __forceinline__ int awg_index(const ChainCtx& c, const int j) {   // DigitalToAnalog, half-pixel nearest
    const int idx = (int)floor((j + 0.5) * c.ratio);
    return idx < c.n_awg - 1 ? idx : c.n_awg - 1;
}

// band-limited I/Q at simulation sample n: resample + direct convolution with the rise function
__device__ __forceinline__ void iq_at(const ChainCtx& c, const double* sI, const double* sQ, const double* sR, const int n,
                                      double& vi, double& vq) {
    if (c.resp_kind == 0) {
        const int idx = awg_index(c, n);
        vi = sI[idx];
        vq = sQ[idx];
        return;
    }
    vi = 0.0; vq = 0.0;
    for (int m = 0; m < c.taps; ++m) {
        const int j = n - c.shift - m;
        if (j < 0) break;
        const int idx = awg_index(c, j);
        vi = fma(sR[m], sI[idx], vi);
        vq = fma(sR[m], sQ[idx], vq);
    }
}

__device__ __forceinline__ double flux_freq(const ChainCtx& c, const double phi) {
    return (c.ch[CH_OMEGA0] - c.ch[CH_ANHAR]) * flux_factor(phi, c.ch[CH_PHI0], c.ch[CH_D], c.has_d) + c.ch[CH_ANHAR];
}

__global__ void __launch_bounds__(128) signal_chain_kernel(const SignalParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sI = reinterpret_cast<double*>(smem_raw);    // [max_awg]
    double* sQ = sI + p.max_awg;                         // [max_awg]
    double* sR = sQ + p.max_awg;                         // [max_taps]
    __shared__ double s_norm;

    const int bk = blockIdx.x;
    const int b = bk / p.K, k = bk - b * p.K;
    const ChainCtx c = chain_ctx(p, bk, k);
    fill_awg(p, c, b, k, sI, sQ);

    // ---- noise devices (c3/generator/devices.py:943-1035) ------------------------------------------------------------
    const double* nz = p.noise ? p.noise + (p.noise_batched ? (size_t)bk : (size_t)k) * NOISE_NPAR : nullptr;
    double* tr = p.noise_out ? p.noise_out + (size_t)bk * NOISE_NTRACE * c.N : nullptr;
    int* sPink = reinterpret_cast<int*>(sR + p.max_taps);          // [N] sum of the bistable fluctuators (Pink_Noise)
    double dc = 0.0;
    int bfl = 0;
    if (nz != nullptr) {
        if (tr != nullptr)
            for (int e = threadIdx.x; e < NOISE_NTRACE * c.N; e += blockDim.x) tr[e] = 0.0;
        // Additive_Noise behind the AWG: independent Gaussian noise on every in-phase and quadrature sample
        if (nz[NOISE_AWG_AMP] >= 1e-17) {
            __syncthreads();
            for (int j = threadIdx.x; j < c.n_awg; j += blockDim.x) {
                double z0, z1;
                noise_normals(p.seed, bk, STREAM_AWG, (unsigned int)j, z0, z1);
                sI[j] += nz[NOISE_AWG_AMP] * z0;
                sQ[j] += nz[NOISE_AWG_AMP] * z1;
                if (tr != nullptr && j < c.N) { tr[TRACE_AWG_I * c.N + j] = nz[NOISE_AWG_AMP] * z0; tr[TRACE_AWG_Q * c.N + j] = nz[NOISE_AWG_AMP] * z1; }
            }
        }
        // DC_Noise: one Gaussian offset per realisation
        if (nz[NOISE_DC_AMP] >= 1e-17) {
            double z0, z1;
            noise_normals(p.seed, bk, STREAM_DC, 0u, z0, z1);
            dc = nz[NOISE_DC_AMP] * z0;
        }
        // Pink_Noise: bfl_num two-level fluctuators, fluctuator i flips at a step with probability 1 / rate_i,
        // rate = logspace(0, ln N, bfl_num + 1, base 10)[1:]  (the reference's np.logspace(0, np.log(num_steps), ...)):
        // one thread per fluctuator walks the time axis, the others wait (N * bfl_num draws: microseconds)
        bfl = (nz[NOISE_PINK_AMP] >= 1e-17) ? min(32, max(0, (int)nz[NOISE_PINK_BFL])) : 0;
        if (bfl > 0) {
            for (int n = threadIdx.x; n < c.N; n += blockDim.x) sPink[n] = 0;
            __syncthreads();
            if ((int)threadIdx.x < bfl) {
                unsigned int w[4];
                noise_words(p.seed, bk, STREAM_PINK_INIT, threadIdx.x, w);
                int state = (w[0] & 1u) ? 1 : -1;
                const double rate = pow(10.0, log((double)c.N) * (double)(threadIdx.x + 1) / (double)bfl);
                for (int n = 0; n < c.N; ++n) {
                    if ((n & 1) == 0) noise_words(p.seed, bk, STREAM_PINK_FLIP, (unsigned int)(threadIdx.x * ((c.N + 1) / 2) + (n >> 1)), w);
                    const double u = (n & 1) ? u01(w[2], w[3]) : u01(w[0], w[1]);
                    if (floor(u * rate) == 0.0) state = -state;
                    atomicAdd(&sPink[n], state);
                }
            }
        }
    }
    fill_taps(c, sR, &s_norm);                                   // (contains the barriers that publish sI / sQ / sPink)

    // ---- simulation grid: resample, convolve, mix with the LO, convert ------------------------------------------
    const double f_ref = (c.out_kind == 1) ? flux_freq(c, c.ch[CH_PHI]) : 0.0;
    double* out = p.out + (size_t)bk * c.N;
    for (int n = threadIdx.x; n < c.N; n += blockDim.x) {
        double vi, vq;
        iq_at(c, sI, sQ, sR, n, vi, vq);
        const double t = linspace_at(c.s0, c.s1, c.N, n);
        double sn, cs;
        sincos(c.w_lo * t, &sn, &cs);
        double extra = 0.0;
        if (nz != nullptr) {
            if (nz[NOISE_LO_PERC] >= 1e-17) {                    // LONoise
                double z0, z1;
                noise_normals(p.seed, bk, STREAM_LO, (unsigned int)n, z0, z1);
                cs += nz[NOISE_LO_PERC] * z0;
                sn += nz[NOISE_LO_PERC] * z1;
                if (tr != nullptr) { tr[TRACE_LO_COS * c.N + n] = nz[NOISE_LO_PERC] * z0; tr[TRACE_LO_SIN * c.N + n] = nz[NOISE_LO_PERC] * z1; }
            }
            if (nz[NOISE_ADD_AMP] >= 1e-17) {                    // Additive_Noise behind the mixer
                double z0, z1;
                noise_normals(p.seed, bk, STREAM_ADD, (unsigned int)n, z0, z1);
                extra += nz[NOISE_ADD_AMP] * z0;
                if (tr != nullptr) tr[TRACE_ADD * c.N + n] = nz[NOISE_ADD_AMP] * z0;
            }
            const double pink = bfl > 0 ? nz[NOISE_PINK_AMP] * (double)sPink[n] : 0.0;
            extra += dc + pink + nz[NOISE_DC_OFFSET];
            if (tr != nullptr) { tr[TRACE_DC * c.N + n] = dc; tr[TRACE_PINK * c.N + n] = pink; }
        }
        const double mixed = cs * vi + sn * vq + extra;
        out[n] = (c.out_kind == 1) ? flux_freq(c, c.ch[CH_PHI] + mixed) - f_ref : mixed * c.ch[CH_V2HZ];
    }
}

// ---- reverse mode: dL/d(pulse parameters) from dL/d(signals) --------------------------------------------------------
// The chain is linear from the AWG samples to the mixer output, so its adjoint is exact and cheap:
//   h[n]   = gsig[n] * d out / d mixed * (cos, sin)(w t_n)
//   a[j]   = sum over the simulation samples j' resampled from AWG sample j of sum_m r[m] h[j' + shift + m]
//   dL/dp  = sum_j a_I[j] dI_j/dp + a_Q[j] dQ_j/dp, with dI/dp, dQ/dp from the dual-number envelope evaluation
// plus the direct terms for the carrier frequency and V_to_Hz.  Replaces tf.GradientTape through
// Generator.generate_signals (the reference's gradient-based optimal control of pulse parameters,
// c3/optimizers/optimalcontrol.py:200-228).
struct SignalGradParams {
    SignalParams f;          // forward description (f.out unused)
    const double* gsig;      // [B, K, N]
    double* genv;            // [B, K, E, ENV_NPAR]
    double* glo;             // [B, K]
    double* gv2hz;           // [B, K] or null
};

__device__ __forceinline__ double block_sum_128(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    return red[0] + red[1] + red[2] + red[3];
}

__global__ void __launch_bounds__(128) signal_chain_grad_kernel(const SignalGradParams g) {
    const SignalParams& p = g.f;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sI = reinterpret_cast<double*>(smem_raw);    // [max_awg] forward AWG samples
    double* sQ = sI + p.max_awg;
    double* aI = sQ + p.max_awg;                         // [max_awg] adjoint of the AWG samples
    double* aQ = aI + p.max_awg;
    double* sR = aQ + p.max_awg;                         // [max_taps]
    double* hI = sR + p.max_taps;                        // [N]
    double* hQ = hI + p.N;
    __shared__ double s_norm;
    __shared__ double red[4];

    const int bk = blockIdx.x;
    const int b = bk / p.K, k = bk - b * p.K;
    const ChainCtx c = chain_ctx(p, bk, k);
    fill_awg(p, c, b, k, sI, sQ);
    fill_taps(c, sR, &s_norm);

    // ---- pass 1 over the simulation grid: direct terms and h -------------------------------------------------------
    const double* gs = g.gsig + (size_t)bk * c.N;
    double acc_lo = 0.0, acc_v = 0.0;
    for (int n = threadIdx.x; n < c.N; n += blockDim.x) {
        double vi, vq;
        iq_at(c, sI, sQ, sR, n, vi, vq);
        const double t = linspace_at(c.s0, c.s1, c.N, n);
        double sn, cs;
        sincos(c.w_lo * t, &sn, &cs);
        const double mixed = cs * vi + sn * vq;
        double dout;                                      // d out / d mixed
        if (c.out_kind == 1) {
            const double x = M_PI * (c.ch[CH_PHI] + mixed) / c.ch[CH_PHI0];
            const double cx = cos(x), sx = sin(x);
            double dfac;
            if (c.has_d) {
                const double d2 = c.ch[CH_D] * c.ch[CH_D];
                const double q = cx * cx + d2 * sx * sx;
                dfac = 0.5 * (d2 - 1.0) * sx * cx / (sqrt(sqrt(q)) * sqrt(q));
            } else {
                dfac = -0.5 * sx * (cx >= 0 ? 1.0 : -1.0) / sqrt(fabs(cx));
            }
            dout = (c.ch[CH_OMEGA0] - c.ch[CH_ANHAR]) * dfac * M_PI / c.ch[CH_PHI0];
        } else {
            dout = c.ch[CH_V2HZ];
            acc_v += gs[n] * mixed;
        }
        const double gg = gs[n] * dout;
        hI[n] = gg * cs;
        hQ[n] = gg * sn;
        acc_lo += gg * t * (-sn * vi + cs * vq);
    }
    const double tot_lo = block_sum_128(acc_lo, red);
    const double tot_v = block_sum_128(acc_v, red);
    if (threadIdx.x == 0) {
        g.glo[bk] = tot_lo;
        if (g.gv2hz) g.gv2hz[bk] = tot_v;
    }
    __syncthreads();

    // ---- pass 2: adjoint of convolution + resampling, one AWG sample per thread (deterministic order) ---------------
    for (int j = threadIdx.x; j < c.n_awg; j += blockDim.x) {
        // simulation samples [lo, hi) that the DigitalToAnalog maps to AWG sample j (the map is monotone)
        int lo = (int)ceil(j / c.ratio - 0.5) - 1;
        if (lo < 0) lo = 0;
        while (lo < c.N && awg_index(c, lo) < j) ++lo;
        while (lo > 0 && awg_index(c, lo - 1) >= j) --lo;
        double ai = 0.0, aq = 0.0;
        for (int jp = lo; jp < c.N && awg_index(c, jp) == j; ++jp) {
            if (c.resp_kind == 0) {
                ai += hI[jp];
                aq += hQ[jp];
            } else {
                for (int m = 0; m < c.taps; ++m) {
                    const int n = jp + c.shift + m;
                    if (n >= c.N) break;
                    ai = fma(sR[m], hI[n], ai);
                    aq = fma(sR[m], hQ[n], aq);
                }
            }
        }
        aI[j] = ai;
        aQ[j] = aq;
    }
    __syncthreads();

    // ---- pass 3: envelope parameter derivatives by dual numbers --------------------------------------------------------
    for (int e = 0; e < p.E; ++e) {
        const int id = p.shape[k * p.E + e];
        double* out = g.genv + (((size_t)b * p.K + k) * p.E + e) * ENV_NPAR;
        if (id < 0) {
            if (threadIdx.x < ENV_NPAR) out[threadIdx.x] = 0.0;
            continue;
        }
        if (id >= SHAPE_FIRST_EXT) {           // extended shapes are forward only: a loud NaN, never a silent zero
            if (threadIdx.x < ENV_NPAR) out[threadIdx.x] = __longlong_as_double(0x7ff8000000000000LL);
            continue;
        }
        const int fl = p.flags[k * p.E + e];
        const double* ev = p.env + (((size_t)b * p.K + k) * p.E + e) * ENV_NPAR;
        for (int q = 0; q < ENV_NPAR; ++q) {
            Dual dv[ENV_NPAR];
#pragma unroll
            for (int i = 0; i < ENV_NPAR; ++i) dv[i] = Dual(ev[i], i == q ? 1.0 : 0.0);
            double acc = 0.0;
            for (int j = threadIdx.x; j < c.n_awg; j += blockDim.x) {
                const double t = linspace_at(c.a0, c.a1, c.n_awg, j) - p.t_start;
                Dual re, im;
                awg_term<Dual>(id, fl, t, dv, c.off0, c.off1, re, im);
                acc += aI[j] * re.d + aQ[j] * im.d;
            }
            const double tot = block_sum_128(acc, red);
            if (threadIdx.x == 0) out[q] = tot;
        }
    }
}

}  // namespace c3b
