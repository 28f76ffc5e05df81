// slic_port.cpp -- TEST INFRASTRUCTURE ONLY (oracle).  Plain restatement of the SLIC stage of the reference's `cluster`
// (/root/reference/src/cluster.cc) in the order-independent form the GPU kernels use, checked against the reference's own code
// (oracle/_ref, ref_slic) by tests/test_oracle_slic.py:
//   SLIC()            :295-344   Lab image in (cvtColor is the input boundary), Sobel x/y in double, 0.5/0.5 blend, 5 rounds
//   initilizeCenters  :207-237   one centre per len x len block at (j + len/2, i + len/2), label = running number from 1
//   fituneCenter      :241-283   centre moves to the 3x3 neighbour with the smallest squared gradient (first minimum in row-major order)
//   clustering        :88-147    a pixel takes the centre with the smallest dis = sqrt(disc^2 + m diss^2) among the centres whose
//                                [c - len, c + len) window holds it; the reference walks the centres in index order with a strict <,
//                                i.e. ties go to the LOWEST index; a pixel no window covers keeps its label of the round before
//   updateCenter      :160-203   centre := truncated means of x, y, L, A, B, D over the pixels of its window that carry its label
// Arithmetic: doubles, every operation individually rounded (the oracle build uses -ffp-contract=off); pow(x, 2) is x * x.
#include <cstdint>
#include <cmath>
#include <cstring>
#include <vector>

namespace {
struct Center { int x, y, L, A, B, D, label; };
inline int refl(int p, int n) { if (n == 1) return 0; while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p; return p; }
}

extern "C" int port_slic(const uint8_t* lab, const uint16_t* depth, int rows, int cols, int len, int m, double* labels, int* centers_out, int cap, int* n_out) {
    auto px = [&](int y, int x, int c) { return (int)lab[((size_t)y * cols + x) * 3 + c]; };
    // gradient: 0.5 * Sobel(dy) + 0.5 * Sobel(dx), per channel (integers and halves: exact)
    auto grad = [&](int y, int x, int c) {
        int gx = 0, gy = 0;
        static const int D[3] = {-1, 0, 1}, S[3] = {1, 2, 1};
        for (int i = -1; i <= 1; ++i) for (int j = -1; j <= 1; ++j) {
            const int v = px(refl(y + i, rows), refl(x + j, cols), c);
            gy += D[i + 1] * S[j + 1] * v; gx += S[i + 1] * D[j + 1] * v;
        }
        return (double)gy * 0.5 + (double)gx * 0.5 + 0.0;
    };
    std::vector<Center> cs;
    int num = 0;
    for (int i = 0; i < rows; i += len) {
        const int cy = i + len / 2; if (cy >= rows) continue;
        for (int j = 0; j < cols; j += len) {
            const int cx = j + len / 2; if (cx >= cols) continue;
            Center c; c.x = cx; c.y = cy; c.L = px(cy, cx, 0); c.A = px(cy, cx, 1); c.B = px(cy, cx, 2); c.label = ++num; c.D = depth[(size_t)cy * cols + cx];
            cs.push_back(c);
        }
    }
    for (Center& c : cs) {
        if (c.x - 1 < 0 || c.x + 1 >= cols || c.y - 1 < 0 || c.y + 1 >= rows) continue;
        double best = 9999999; int tx = 0, ty = 0;
        for (int mm = -1; mm < 2; ++mm) for (int nn = -1; nn < 2; ++nn) {
            const double g0 = grad(c.y + mm, c.x + nn, 0), g1 = grad(c.y + mm, c.x + nn, 1), g2 = grad(c.y + mm, c.x + nn, 2);
            const double g = g0 * g0 + g1 * g1 + g2 * g2;
            if (g < best) { best = g; ty = mm; tx = nn; }
        }
        c.x += tx; c.y += ty; c.L = px(c.y, c.x, 0); c.A = px(c.y, c.x, 1); c.B = px(c.y, c.x, 2);
    }
    std::vector<double> lbl((size_t)rows * cols, 0.0);
    for (int round = 0; round < 5; ++round) {
        // assignment, pixel by pixel: minimum over the covering centres, ties to the lowest index
        for (int y = 0; y < rows; ++y) for (int x = 0; x < cols; ++x) {
            double best = 999999; int who = -1;
            const int L = px(y, x, 0), A = px(y, x, 1), B = px(y, x, 2);
            for (size_t k = 0; k < cs.size(); ++k) {
                const Center& c = cs[k];
                if (x < c.x - len || x >= c.x + len || y < c.y - len || y >= c.y + len) continue;
                const double disc = std::sqrt((double)((L - c.L) * (L - c.L)) + (double)((A - c.A) * (A - c.A)) + (double)((B - c.B) * (B - c.B)));
                const double diss = std::sqrt((double)((x - c.x) * (x - c.x)) + (double)((y - c.y) * (y - c.y)));
                const double dis = std::sqrt(disc * disc + (double)m * (diss * diss));
                if (dis < best) { best = dis; who = (int)k; }
            }
            if (who >= 0) lbl[(size_t)y * cols + x] = cs[who].label;
        }
        for (Center& c : cs) {
            double sx = 0, sy = 0, sL = 0, sA = 0, sB = 0, sn = 0, sD = 0;
            for (int i = c.y - len; i < c.y + len; ++i) { if (i < 0 || i >= rows) continue;
                for (int j = c.x - len; j < c.x + len; ++j) { if (j < 0 || j >= cols) continue;
                    if (lbl[(size_t)i * cols + j] == c.label) { sL += px(i, j, 0); sA += px(i, j, 1); sB += px(i, j, 2); sx += j; sy += i; sn += 1; sD += depth[(size_t)i * cols + j]; } } }
            if (sn == 0) sn = 0.000000001;
            c.x = (int)(sx / sn); c.y = (int)(sy / sn); c.L = (int)(sL / sn); c.A = (int)(sA / sn); c.B = (int)(sB / sn); c.D = (int)(sD / sn);
        }
    }
    std::memcpy(labels, lbl.data(), sizeof(double) * lbl.size());
    *n_out = (int)cs.size();
    for (int i = 0; i < (int)cs.size() && i < cap; ++i) { int* o = centers_out + (size_t)i * 7; o[0] = cs[i].x; o[1] = cs[i].y; o[2] = cs[i].L; o[3] = cs[i].A; o[4] = cs[i].B; o[5] = cs[i].D; o[6] = cs[i].label; }
    return 0;
}
