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
enum { SHAPE_NO_DRIVE = 0, SHAPE_RECT, SHAPE_GAUSSIAN_NONORM, SHAPE_GAUSSIAN_SIGMA, SHAPE_COSINE, SHAPE_FLATTOP };

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
};

__device__ __forceinline__ double sigmoid_(double x) { return 1.0 / (1.0 + exp(-x)); }

__device__ __forceinline__ double shape_value(int id, double t, const double* e) {
    switch (id) {
        case SHAPE_RECT: return 1.0;
        case SHAPE_GAUSSIAN_NONORM: {
            const double u = t - e[ENV_TFINAL] / 2, s = e[ENV_SIGMA];
            return exp(-(u * u) / (2 * s * s));
        }
        case SHAPE_GAUSSIAN_SIGMA: {
            const double tf = e[ENV_TFINAL], s = e[ENV_SIGMA], u = t - tf / 2;
            const double gauss = exp(-(u * u) / (2 * s * s));
            const double offset = exp(-(tf * tf) / (8 * s * s));
            const double norm = sqrt(2 * M_PI * s * s) * erf(tf / (sqrt(8.0) * s)) - tf * offset;
            return (gauss - offset) / norm;
        }
        case SHAPE_COSINE: return 0.5 * (1 - cos(2 * M_PI * t / e[ENV_TFINAL]));
        case SHAPE_FLATTOP:
            return (1 + erf((t - e[ENV_TUP]) / e[ENV_RISEFALL])) / 2 * (1 + erf((-t + e[ENV_TDOWN]) / e[ENV_RISEFALL])) / 2;
        default: return 0.0;
    }
}

__device__ __forceinline__ double shape_deriv(int id, double t, const double* e) {
    switch (id) {
        case SHAPE_GAUSSIAN_NONORM:
        case SHAPE_GAUSSIAN_SIGMA: {
            const double tf = e[ENV_TFINAL], s = e[ENV_SIGMA], u = t - tf / 2;
            double g = exp(-(u * u) / (2 * s * s)) * (-u / (s * s));
            if (id == SHAPE_GAUSSIAN_SIGMA) {
                const double offset = exp(-(tf * tf) / (8 * s * s));
                g /= sqrt(2 * M_PI * s * s) * erf(tf / (sqrt(8.0) * s)) - tf * offset;
            }
            return g;
        }
        case SHAPE_COSINE: return 0.5 * sin(2 * M_PI * t / e[ENV_TFINAL]) * 2 * M_PI / e[ENV_TFINAL];
        case SHAPE_FLATTOP: {
            const double rf = e[ENV_RISEFALL], up = (t - e[ENV_TUP]) / rf, dn = (-t + e[ENV_TDOWN]) / rf;
            const double c = 2 / sqrt(M_PI) / rf;
            return (c * exp(-up * up) * (1 + erf(dn)) - (1 + erf(up)) * c * exp(-dn * dn)) / 4;
        }
        default: return 0.0;
    }
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

__global__ void __launch_bounds__(128) signal_chain_kernel(const SignalParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sI = reinterpret_cast<double*>(smem_raw);    // [max_awg]
    double* sQ = sI + p.max_awg;                         // [max_awg]
    double* sR = sQ + p.max_awg;                         // [max_taps]
    __shared__ double s_norm;

    const int bk = blockIdx.x;
    const int b = bk / p.K, k = bk - b * p.K;
    const double* ch = p.chain + (p.chain_batched ? (size_t)bk : (size_t)k) * CH_NPAR;
    const double sim_res = ch[CH_SIM_RES], awg_res = ch[CH_AWG_RES], rise = ch[CH_RISE_TIME];
    const int resp_kind = (int)ch[CH_RESP_KIND], out_kind = (int)ch[CH_OUT_KIND];
    const double span = fabs(p.t_start - p.t_end);
    const int n_awg = (int)(span * awg_res);
    const int N = p.N;
    const double dt_awg = 1.0 / awg_res, dt_sim = 1.0 / sim_res;
    const double a0 = p.t_start + dt_awg / 2, a1 = p.t_end - dt_awg / 2;

    // ---- AWG samples: sum of the line's envelopes on the AWG time grid -------------------------------------
    const double step = n_awg > 1 ? linspace_at(a0, a1, n_awg, 1) - linspace_at(a0, a1, n_awg, 0) : dt_awg;  // ts[1] - ts[0]
    const double off0 = linspace_at(a0, a1, n_awg, 0) - p.t_start;
    const double off1 = n_awg > 1 ? linspace_at(a0, a1, n_awg, 1) - p.t_start : off0 + dt_awg;
    const double dts = off1 - off0;                      // ts_off[1] - ts_off[0]
    for (int j = threadIdx.x; j < n_awg; j += blockDim.x) {
        const double t = linspace_at(a0, a1, n_awg, j) - p.t_start;
        double re = 0.0, im = 0.0;
        for (int e = 0; e < p.E; ++e) {
            const int id = p.shape[k * p.E + e];
            if (id < 0) continue;
            const int fl = p.flags[k * p.E + e];
            const double* ev = p.env + (((size_t)b * p.K + k) * p.E + e) * ENV_NPAR;
            const double tfin = ev[ENV_TFINAL];
            const double mask = sigmoid_((t / dts + 0.001) * 1e6) * sigmoid_((0.999 * tfin - t) / dts * 1e6);
            double sv = shape_value(id, t, ev);
            if (fl & 2) sv -= shape_value(id, 2 * off0 - off1, ev);
            const double env_re = mask * sv;
            const double env_im = (fl & 1) ? -(mask * shape_deriv(id, t, ev) * dts) * ev[ENV_DELTA] : 0.0;
            double sn, cs;
            sincos(ev[ENV_XY] - ev[ENV_FREQ_OFFSET] * t, &sn, &cs);
            const double amp = ev[ENV_AMP];
            re += amp * (env_re * cs - env_im * sn);
            im += amp * (env_re * sn + env_im * cs);
        }
        sI[j] = re;
        sQ[j] = im;
    }
    (void)step;

    // ---- Gaussian rise function of the Response device ---------------------------------------------------------
    int taps = 0;
    if (resp_kind != 0) {
        taps = (int)floor(rise * sim_res);
        const double cen = (resp_kind == 2) ? (rise - 1 / sim_res) / 2 : (rise + 1 / sim_res) / 2;
        const double sg = rise / 4;
        const double offset = exp(-((-1 - cen) * (-1 - cen)) / (2 * sg * sg));
        for (int m = threadIdx.x; m < taps; m += blockDim.x) {
            const double tr = linspace_at(0.0, rise, taps, m);
            sR[m] = exp(-((tr - cen) * (tr - cen)) / (2 * sg * sg)) - offset;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int m = 0; m < taps; ++m) s += sR[m];
            s_norm = s;
        }
        __syncthreads();
        for (int m = threadIdx.x; m < taps; m += blockDim.x) sR[m] = sR[m] / s_norm;
    }
    __syncthreads();

    // ---- simulation grid: resample, convolve, mix with the LO, convert ------------------------------------------
    const double s0 = p.t_start + dt_sim / 2, s1 = p.t_end - dt_sim / 2;
    const double ratio = (double)n_awg / (double)N;
    const double w_lo = p.lo_freq[bk];
    const int shift = (resp_kind == 1) ? 1 : 0;
    const bool has_d = ch[CH_D] == ch[CH_D];             // NaN: symmetric SQUID
    double f_ref = 0.0;
    if (out_kind == 1) f_ref = (ch[CH_OMEGA0] - ch[CH_ANHAR]) * flux_factor(ch[CH_PHI], ch[CH_PHI0], ch[CH_D], has_d) + ch[CH_ANHAR];
    double* out = p.out + (size_t)bk * N;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        double vi, vq;
        if (resp_kind == 0) {
            int idx = (int)floor((n + 0.5) * ratio);
            idx = idx < n_awg - 1 ? idx : n_awg - 1;
            vi = sI[idx];
            vq = sQ[idx];
        } else {
            vi = 0.0; vq = 0.0;
            for (int m = 0; m < taps; ++m) {
                const int j = n - shift - m;
                if (j < 0) break;
                int idx = (int)floor((j + 0.5) * ratio);
                idx = idx < n_awg - 1 ? idx : n_awg - 1;
                const double r = sR[m];
                vi = fma(r, sI[idx], vi);
                vq = fma(r, sQ[idx], vq);
            }
        }
        const double t = linspace_at(s0, s1, N, n);
        double sn, cs;
        sincos(w_lo * t, &sn, &cs);
        const double mixed = cs * vi + sn * vq;
        double v;
        if (out_kind == 1) {
            v = (ch[CH_OMEGA0] - ch[CH_ANHAR]) * flux_factor(ch[CH_PHI] + mixed, ch[CH_PHI0], ch[CH_D], has_d) + ch[CH_ANHAR] - f_ref;
        } else {
            v = mixed * ch[CH_V2HZ];
        }
        out[n] = v;
    }
}

}  // namespace c3b
