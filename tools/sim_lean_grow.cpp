// tools/sim_lean_grow.cpp — design check of the LEAN region growing of csrc/line.cu (l_region_grow_lean): a lane-level
// emulation of the warp algorithm (rounds, exact float-degree test with double fallback, robust multi-accept under the drift
// bound, duplicate handling) run in lockstep with the oracle's sequential region_grow on real frames: every call (first growth
// and refine's re-growth with tau) must give the same list in the same order and the same region angle.
// Test infrastructure (includes the oracle source).  Build:
//   g++ -O2 -ffp-contract=off -I oracle tools/sim_lean_grow.cpp oracle/orb_oracle.cpp -o /tmp/sim_lean_grow
//   /tmp/sim_lean_grow scaled.raw w h
#include "../oracle/line_oracle.cpp"
#include <cstdio>
namespace {
struct Stats { long calls = 0, steps = 0, rounds = 0, atan_chain = 0, near = 0, accepted = 0, multi = 0; } ST;
struct Lean {
    const Lsd& L; std::vector<float> angdeg;
    Lean(const Lsd& l, const uchar* img, int pitch, double threshold) : L(l), angdeg((size_t)l.w * l.h, -1024.f) {
        for (int y = 0; y < l.h - 1; y++) for (int x = 0; x < l.w - 1; x++) {
            const uchar* p = img + (size_t)y * pitch + x; int DA = p[pitch + 1] - p[0], BC = p[1] - p[pitch]; int gx = DA + BC, gy = DA - BC;
            double norm = std::sqrt((gx * gx + gy * gy) / 4.0);
            if (norm > threshold) angdeg[(size_t)y * l.w + x] = orc_fast_atan2((float)gx, (float)-gy);
        }
    }
    static bool aligned_rad(double a, double theta, double prec) { double n = theta - a; if (n < 0) n = -n; if (n > M_3_2_PI) { n -= M_2__PI; if (n < 0) n = -n; } return n <= prec; }
    // `used` is the caller's map (modified like the GPU does)
    int grow(std::vector<uchar>& used, int sx, int sy, double prec, std::vector<RegPt>& reg, double& reg_angle) {
        const int w = L.w, h = L.h; ST.calls++;
        reg.clear(); reg.push_back(RegPt{sx, sy}); used[(size_t)sy * w + sx] = 1;
        float th = angdeg[(size_t)sy * w + sx];
        const double ra0 = (double)th * DEG_TO_RADS;
        float sumdx = float(std::cos(ra0)), sumdy = float(std::sin(ra0));
        float rM = 1.0f / std::sqrt(sumdx * sumdx + sumdy * sumdy); bool dirty = false;
        const float pdeg = (float)(prec * (180.0 / PI)); const bool rob_ok = prec < 0.78;
        const float coef = (float)(57.2958 * 1.0002 * std::sin(prec + 0.0006));      // deg per (accepted vector / |S|): |turn| <= sum |sin phi_i| / |S|, |phi_i| <= prec + E
        for (size_t i = 0; i < reg.size();) {
            const int cnt = (int)std::min<size_t>(4, reg.size() - i); ST.steps++;
            int q[32], xx[32], yy[32]; bool valid[32]; float a[32], cx[32], cy[32]; bool cand[32];
            for (int lane = 0; lane < 32; lane++) {
                const int slot = lane >> 3, nb = (lane & 7) + ((lane & 7) >= 4 ? 1 : 0), ox = nb % 3 - 1, oy = nb / 3 - 1;
                valid[lane] = false; cand[lane] = false; q[lane] = -1 - lane;
                if (slot >= cnt) continue;
                xx[lane] = reg[i + slot].x + ox; yy[lane] = reg[i + slot].y + oy;
                if (xx[lane] < 0 || yy[lane] < 0 || xx[lane] >= w || yy[lane] >= h) continue;
                valid[lane] = true; q[lane] = yy[lane] * w + xx[lane];
                a[lane] = angdeg[q[lane]];
                const float ar = float((double)a[lane] * DEG_TO_RADS);
                cx[lane] = cosf_c(ar); cy[lane] = sinf_c(ar);
                cand[lane] = a[lane] != -1024.f && used[q[lane]] == 0;
            }
            if (dirty) { th = orc_fast_atan2(sumdy, sumdx); rM = 1.0f / std::sqrt(sumdx * sumdx + sumdy * sumdy); dirty = false; }
            unsigned pending = 0; for (int l = 0; l < 32; l++) if (cand[l]) pending |= 1u << l;
            while (pending) {
                ST.rounds++;
                unsigned P = 0; float e[32]; bool pass[32];
                for (int l = 0; l < 32; l++) { pass[l] = false; e[l] = 0; if (!((pending >> l) & 1)) continue;
                    const float d = std::fabs(th - a[l]); e[l] = d > 270.f ? 360.f - d : d; pass[l] = e[l] <= pdeg;
                    const bool near = std::fabs(e[l] - pdeg) < 2e-3f || std::fabs(d - 270.f) < 2e-3f;
                    if (near) { ST.near++; pass[l] = aligned_rad((double)a[l] * DEG_TO_RADS, (double)th * DEG_TO_RADS, prec); }
                    if (pass[l]) P |= 1u << l; }
                if (!P) break;
                const int k0 = __builtin_ctz(P);
                unsigned P1 = 0;                                  // first holders among the passing lanes
                for (int l = 0; l < 32; l++) if ((P >> l) & 1) { bool dup = false; for (int m = 0; m < l; m++) if (((P >> m) & 1) && q[m] == q[l]) dup = true; if (!dup) P1 |= 1u << l; }
                unsigned NR = 0;
                for (int l = k0 + 1; l < 32; l++) { if (!((pending >> l) & 1)) continue; bool robust = false;
                    const float x = (float)__builtin_popcount(P1 & ((1u << l) - 1u)) * rM;
                    if (rob_ok && x <= 0.5f) { const float B = coef * x + 0.0215f; robust = pass[l] ? (e[l] <= pdeg - B) : (e[l] >= pdeg + B); }
                    if (!robust) NR |= 1u << l; }
                const unsigned below = NR ? ((NR & (0u - NR)) - 1u) : 0xffffffffu;
                const unsigned Ac = P & below; unsigned A = 0;
                for (int l = 0; l < 32; l++) if ((Ac >> l) & 1) { bool dup = false; for (int m = 0; m < l; m++) if (((Ac >> m) & 1) && q[m] == q[l]) dup = true; if (!dup) A |= 1u << l; }
                if (__builtin_popcount(A) > 1) ST.multi++;
                for (int l = 0; l < 32; l++) if ((A >> l) & 1) { sumdx += cx[l]; sumdy += cy[l]; reg.push_back(RegPt{xx[l], yy[l]}); used[q[l]] = 1; ST.accepted++; }
                pending &= ~below;
                for (int l = 0; l < 32; l++) if ((pending >> l) & 1) for (int m = 0; m < 32; m++) if (((A >> m) & 1) && q[m] == q[l]) pending &= ~(1u << l);
                dirty = true;
                if (pending) { ST.atan_chain++; th = orc_fast_atan2(sumdy, sumdx); rM = 1.0f / std::sqrt(sumdx * sumdx + sumdy * sumdy); dirty = false; }
            }
            i += cnt;
        }
        if (dirty) th = orc_fast_atan2(sumdy, sumdx);
        reg_angle = (double)th * DEG_TO_RADS;
        return (int)reg.size();
    }
};
}
int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: sim_lean_grow scaled_raw_u8 w h\n"); return 1; }
    const int w = atoi(argv[2]), h = atoi(argv[3]);
    std::vector<uchar> img((size_t)w * h); FILE* f = fopen(argv[1], "rb"); if (!f || fread(img.data(), 1, img.size(), f) != img.size()) return 1; fclose(f);
    Lsd lsd; lsd.w = w; lsd.h = h;
    const double ANG_TH = 22.5, prec = PI * ANG_TH / 180, p = ANG_TH / 180, rho = 2.0 / std::sin(prec);
    lsd.ll_angle(img.data(), w, rho, 1024);
    lsd.LOG_NT = 5 * (std::log10(double(w)) + std::log10(double(h))) / 2 + std::log10(11.0);
    const size_t mrs = size_t(-lsd.LOG_NT / std::log10(p));
    lsd.used.assign((size_t)w * h, 0);
    Lean lean(lsd, img.data(), w, rho);
    // check angdeg against the oracle's double angles
    for (size_t i = 0; i < lsd.angles.size(); i++) { const double a = lean.angdeg[i] == -1024.f ? NOTDEF : (double)lean.angdeg[i] * DEG_TO_RADS; if (a != lsd.angles[i]) { printf("angle map differs at %zu\n", i); return 2; } }
    long bad = 0, regions = 0;
    // hook: run the oracle's loop, but every region_grow call (incl. the one inside refine) is shadowed by the lean emulation
    struct Shadow : Lsd { Lean* lean; long* bad; std::vector<RegPt> r2;
        void check_grow(int sx, int sy, std::vector<RegPt>& reg, double& reg_angle, double prec) {
            std::vector<uchar> u2 = used; double ra2;
            region_grow(sx, sy, reg, reg_angle, prec);
            lean->grow(u2, sx, sy, prec, r2, ra2);
            bool same = r2.size() == reg.size() && ra2 == reg_angle && u2 == used;
            if (same) for (size_t i = 0; i < reg.size(); i++) if (reg[i].x != r2[i].x || reg[i].y != r2[i].y) { same = false; break; }
            if (!same) { (*bad)++; if (*bad < 10) printf("MISMATCH seed (%d,%d) prec %.6f: ref n=%zu angle %.17g, lean n=%zu angle %.17g\n", sx, sy, prec, reg.size(), reg_angle, r2.size(), ra2); }
        }
    } S; S.w = w; S.h = h; S.angles = lsd.angles; S.modgrad = lsd.modgrad; S.order = lsd.order; S.LOG_NT = lsd.LOG_NT; S.used.assign((size_t)w * h, 0); S.lean = &lean; S.bad = &bad;
    Lean lean2(S, img.data(), w, rho); S.lean = &lean2;
    std::vector<RegPt> reg;
    for (int idx : S.order) {
        if (S.used[idx] != 0 || S.angles[idx] == NOTDEF) continue;
        double reg_angle; regions++;
        S.check_grow(idx % w, idx / w, reg, reg_angle, prec);
        if (reg.size() < mrs) continue;
        Rect rec; S.region2rect(reg, reg_angle, prec, p, rec);
        // refine, with its region_grow shadowed (restated from Lsd::refine)
        double density = double(reg.size()) / (Lsd::dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
        if (density >= 0.7) continue;
        const double xc = double(reg[0].x), yc = double(reg[0].y), ang_c = S.angles[(size_t)reg[0].y * w + reg[0].x];
        double sum = 0, s_sum = 0; int n = 0;
        for (const RegPt& r : reg) { S.used[(size_t)r.y * w + r.x] = 0; if (Lsd::dist(xc, yc, double(r.x), double(r.y)) < rec.width) { const double d = Lsd::angle_diff_signed(S.angles[(size_t)r.y * w + r.x], ang_c); sum += d; s_sum += d * d; ++n; } }
        const double mean_angle = sum / double(n), tau = 2.0 * std::sqrt((s_sum - 2.0 * mean_angle * sum) / double(n) + mean_angle * mean_angle);
        const int sx = reg[0].x, sy = reg[0].y;
        S.check_grow(sx, sy, reg, reg_angle, tau);
        if (reg.size() < 2) continue;
        S.region2rect(reg, reg_angle, prec, p, rec);
        density = double(reg.size()) / (Lsd::dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
        if (density < 0.7) S.reduce_region_radius(reg, reg_angle, prec, p, rec, density, 0.7);
    }
    printf("%s %dx%d: %ld seeds grown, %ld grow calls, %ld mismatches | steps %ld rounds %ld (%.2f/step) atan-on-chain %ld multi-accept rounds %ld near-threshold double tests %ld accepted %ld\n",
           argv[1], w, h, regions, ST.calls, bad, ST.steps, ST.rounds, double(ST.rounds) / ST.steps, ST.atan_chain, ST.multi, ST.near, ST.accepted);
    return bad ? 3 : 0;
}
