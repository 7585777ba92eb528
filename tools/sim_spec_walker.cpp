// tools/sim_spec_walker.cpp — design check for the multi-warp LSD region walker (csrc/line.cu, k_lsd_regions).
// Simulates the protocol on real frames with the oracle's own LSD routines (test infrastructure; includes the oracle source):
//   * the ordered seed list is cut into chunks of C seeds; W workers take chunks in order and run them SPECULATIVELY against
//     the committed `used` map as it is when they start; chunks commit strictly in order;
//   * at its turn a chunk is valid iff none of the pixels it ever accepted has been committed by an earlier chunk meanwhile
//     (then every membership test it made had the outcome of the sequential algorithm); otherwise it is redone on the spot.
// Reports: validity rate, the final segments compared with the sequential detector (must be identical), and the critical-path
// speed-up under a simple cost model (cycles per 4-entry grow step, per region, per chunk hand-off).
// Build: g++ -O2 -ffp-contract=off -I oracle tools/sim_spec_walker.cpp oracle/orb_oracle.cpp -o /tmp/sim_spec_walker
#include "../oracle/line_oracle.cpp"
#include <deque>
#include <cstdio>

namespace {
struct Cost { double step = 1500, per_px_rect = 60, region = 4000, chunk = 800, handoff = 300, validate_px = 4; };

struct ChunkResult { std::vector<int> accepted; std::vector<int> finalpix; std::vector<Rect> recs; double cost = 0; };

// one chunk of seeds processed sequentially on `lsd.used` (which the caller has set to the snapshot)
void run_chunk(Lsd& lsd, const std::vector<int>& seeds, size_t a, size_t b, const std::vector<uchar>& base, ChunkResult& R, const Cost& C,
               double prec, double p, size_t min_reg_size) {
    R.accepted.clear(); R.finalpix.clear(); R.recs.clear(); R.cost = C.chunk;
    std::vector<RegPt> reg;
    for (size_t s = a; s < b; s++) {
        const int idx = seeds[s];
        if (lsd.used[idx]) continue;
        double reg_angle;
        lsd.region_grow(idx % lsd.w, idx / lsd.w, reg, reg_angle, prec);
        for (auto& r : reg) R.accepted.push_back(r.y * lsd.w + r.x);
        R.cost += std::ceil(reg.size() / 4.0) * C.step + 500;
        if (reg.size() < min_reg_size) continue;
        Rect rec;
        lsd.region2rect(reg, reg_angle, prec, p, rec);
        R.cost += C.region + reg.size() * C.per_px_rect;
        const size_t n0 = reg.size();
        const double density = double(reg.size()) / (Lsd::dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
        const bool ok = lsd.refine(reg, reg_angle, prec, p, rec, 0.7);
        if (density < 0.7) { for (auto& r : reg) R.accepted.push_back(r.y * lsd.w + r.x); R.cost += std::ceil(reg.size() / 4.0) * C.step + (n0 + reg.size()) * C.per_px_rect; }
        if (!ok) continue;
        R.recs.push_back(rec);
    }
    for (size_t i = 0; i < lsd.used.size(); i++) if (lsd.used[i] && !base[i]) R.finalpix.push_back((int)i);
}
}  // namespace

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: sim_spec_walker raw_u8_file w h [W=8] [C=32]\n"); return 1; }
    const int w = atoi(argv[2]), h = atoi(argv[3]);
    const int W = argc > 4 ? atoi(argv[4]) : 8, CH = argc > 5 ? atoi(argv[5]) : 32;
    std::vector<uchar> img((size_t)w * h);
    FILE* f = fopen(argv[1], "rb"); if (!f || fread(img.data(), 1, img.size(), f) != img.size()) { fprintf(stderr, "read failed\n"); return 1; } fclose(f);
    Cost C;
    Lsd lsd; lsd.w = w; lsd.h = h;
    const double ANG_TH = 22.5, prec = PI * ANG_TH / 180, p = ANG_TH / 180, rho = 2.0 / std::sin(prec);
    lsd.ll_angle(img.data(), w, rho, 1024);
    lsd.LOG_NT = 5 * (std::log10(double(w)) + std::log10(double(h))) / 2 + std::log10(11.0);
    const size_t min_reg_size = size_t(-lsd.LOG_NT / std::log10(p));
    std::vector<int> seeds;
    for (int idx : lsd.order) if (lsd.angles[idx] != NOTDEF) seeds.push_back(idx);
    const size_t nchunks = (seeds.size() + CH - 1) / CH;
    // sequential reference
    std::vector<Rect> seq_recs; double seq_cost = 0;
    {
        lsd.used.assign((size_t)w * h, 0);
        std::vector<uchar> base(lsd.used);
        ChunkResult R;
        for (size_t k = 0; k < nchunks; k++) { base = lsd.used; run_chunk(lsd, seeds, k * CH, std::min(seeds.size(), (k + 1) * CH), base, R, C, prec, p, min_reg_size); seq_cost += R.cost - C.chunk + 100; for (auto& r : R.recs) seq_recs.push_back(r); }
    }
    // speculative pipeline: versions[k] = committed map after chunk k-1 (versions[0] = empty)
    std::deque<std::vector<uchar>> versions; versions.push_back(std::vector<uchar>((size_t)w * h, 0));
    size_t first_version = 0;                       // chunk index of versions.front()
    std::vector<double> commit_end(nchunks + 1, 0.0), worker_free(W, 0.0);
    std::vector<Rect> recs; size_t nvalid = 0, nwork = 0, nredo = 0; double redo_cost = 0;
    for (size_t k = 0; k < nchunks; k++) {
        int wk = 0; for (int i = 1; i < W; i++) if (worker_free[i] < worker_free[wk]) wk = i;
        const double t_claim = worker_free[wk];
        // snapshot: the latest committed version whose commit finished by t_claim
        size_t j = k;                                // version index = number of chunks committed
        while (j > first_version && commit_end[j] > t_claim) j--;
        lsd.used = versions[j - first_version];
        const std::vector<uchar> snap = lsd.used;
        ChunkResult R;
        run_chunk(lsd, seeds, k * CH, std::min(seeds.size(), (k + 1) * CH), snap, R, C, prec, p, min_reg_size);
        const double t_spec_end = t_claim + R.cost;
        const double t_turn = std::max(t_spec_end, commit_end[k]) + C.handoff;
        const std::vector<uchar>& truth = versions[k - first_version];
        bool valid = true;
        for (int q : R.accepted) if (truth[q]) { valid = false; break; }
        double t_end = t_turn + R.accepted.size() * C.validate_px / 32.0 * 8;
        if (!R.accepted.empty()) nwork++;
        if (!valid) {
            lsd.used = truth;
            run_chunk(lsd, seeds, k * CH, std::min(seeds.size(), (k + 1) * CH), truth, R, C, prec, p, min_reg_size);
            t_end += R.cost; nredo++; redo_cost += R.cost;
        } else if (!R.accepted.empty()) nvalid++;
        std::vector<uchar> next = truth;
        for (int q : R.finalpix) next[q] = 1;
        for (auto& r : R.recs) recs.push_back(r);
        versions.push_back(std::move(next));
        while (versions.size() > (size_t)W + 2) { versions.pop_front(); first_version++; }
        commit_end[k + 1] = t_end; worker_free[wk] = t_end;
    }
    bool same = recs.size() == seq_recs.size();
    for (size_t i = 0; same && i < recs.size(); i++) same = recs[i].x1 == seq_recs[i].x1 && recs[i].y1 == seq_recs[i].y1 && recs[i].x2 == seq_recs[i].x2 && recs[i].y2 == seq_recs[i].y2 && recs[i].width == seq_recs[i].width;
    printf("seeds %zu chunks %zu (C=%d) W=%d: working chunks %zu valid %zu redone %zu | regions %zu identical=%d | seq %.2f Mcyc, pipeline %.2f Mcyc, speed-up %.2fx (redo share %.1f%%)\n",
           seeds.size(), nchunks, CH, W, nwork, nvalid, nredo, recs.size(), (int)same, seq_cost / 1e6, commit_end[nchunks] / 1e6, seq_cost / commit_end[nchunks], 100 * redo_cost / seq_cost);
    return same ? 0 : 2;
}
