// orbx_extractor_taps.inl -- stage taps for the parity tests (declared in include/orbx_b200_testtaps.h, not part of the product ABI);
// part of orbx_extractor.cu

// DistributeOctTree stage tap: the pipeline's own sort + tree kernels on caller-provided candidates
extern "C" int orbx_debug_distribute(orbx_extractor* h, const orbx_keypoint* cand, int ncand, int minX, int maxX, int minY, int maxY, int N,
                                     orbx_keypoint* out, int cap, int* n_out) {
    if (!h || !n_out || ncand < 0 || (ncand && !cand) || maxX <= minX || maxY <= minY || N < 0) FAIL(ORBX_E_INVALID, "bad arguments");
    *n_out = 0;
    if (ncand == 0) return ORBX_OK;
    if (maxX - minX > ORBX_MAX_DIM || maxY - minY > ORBX_MAX_DIM || ncand >= (1 << 20)) FAIL(ORBX_E_INVALID, "too large");
    CU_TRY(cudaSetDevice(h->device));
    LevelGeom g; std::memset(&g, 0, sizeof(g));
    g.minBX = minX; g.maxBX = maxX; g.minBY = minY; g.maxBY = maxY; g.N = N;
    g.nIni = (int)std::round(static_cast<float>(maxX - minX) / (maxY - minY));
    if (g.nIni < 1 || g.nIni > 15) FAIL(ORBX_E_INVALID, "unsupported aspect ratio");
    g.hX = static_cast<float>(maxX - minX) / g.nIni;
    const int CH = 1024;
    const int nc = (ncand + CH - 1) / CH;
    g.cell_begin = 0; g.cell_count = nc; g.cand_off = 0; g.cand_cap = ncand; g.kp_off = 0; g.kp_cap = std::max(N + 2, 4 * g.nIni) + 2;
    const int tcap = g.kp_cap + 8;
    if (tcap > 32000) FAIL(ORBX_E_INVALID, "N too large");
    std::vector<CellDesc> cells(nc); std::vector<uint16_t> counts(nc); std::vector<uint32_t> packed(ncand);
    for (int c = 0; c < nc; ++c) { std::memset(&cells[c], 0, sizeof(CellDesc)); cells[c].slot = c * CH; counts[c] = (uint16_t)std::min(CH, ncand - c * CH); }
    for (int i = 0; i < ncand; ++i) {
        const int x = (int)cand[i].x, y = (int)cand[i].y, r = (int)cand[i].response;
        if (x < 0 || y < 0 || x > ORBX_MAX_DIM || y > ORBX_MAX_DIM || r < 0 || r > 255 || (float)x != cand[i].x || (float)y != cand[i].y)
            FAIL(ORBX_E_INVALID, "candidates must have integer coordinates in [0,4095] and response in [0,255]");
        packed[i] = (uint32_t)x | ((uint32_t)y << 12) | ((uint32_t)r << 24);
    }
    DevBuf<LevelGeom> dl; DevBuf<CellDesc> dc; DevBuf<uint16_t> dcnt; DevBuf<uint32_t> dslots, doc, dspk, dkp; DevBuf<unsigned long long> dsk; DevBuf<int> dn, dkc;
    struct Guard { DevBuf<LevelGeom>& a; DevBuf<CellDesc>& b; DevBuf<uint16_t>& c; DevBuf<uint32_t>&d, &e, &f, &g; DevBuf<unsigned long long>& hh; DevBuf<int>&i, &j;
                   ~Guard() { a.release(); b.release(); c.release(); d.release(); e.release(); f.release(); g.release(); hh.release(); i.release(); j.release(); } } guard{dl, dc, dcnt, dslots, doc, dspk, dkp, dsk, dn, dkc};
    if (dl.ensure(1) || dc.ensure(nc) || dcnt.ensure(nc) || dslots.ensure(ncand) || doc.ensure(ncand) || dspk.ensure(ncand) || dsk.ensure(ncand) ||
        dkp.ensure(g.kp_cap) || dn.ensure(1) || dkc.ensure(1) || h->d_overflow.ensure(4)) return ORBX_E_CUDA;
    cudaStream_t s = h->stream;
    DevBuf<uint32_t> dtab; struct TabGuard { DevBuf<uint32_t>& t; ~TabGuard() { t.release(); } } tab_guard{dtab};
    { const char* e = std::getenv("ORBX_QT_TABLES");                           // 1: path codes from tables (as the pipeline does), default: computed
      if (e && std::atoi(e) != 0) {
          std::vector<uint32_t> ct; g.code_nx = ORBX_MAX_DIM + 1; g.code_ny = ORBX_MAX_DIM + 1;
          octree_code_tables(g, g.code_nx, g.code_ny, ct);
          if (dtab.ensure(ct.size())) return ORBX_E_CUDA;
          CU_TRY(cudaMemcpyAsync(dtab.p, ct.data(), ct.size() * 4, cudaMemcpyHostToDevice, s));
          CU_TRY(cudaStreamSynchronize(s));
          g.code_x = dtab.p; g.code_y = dtab.p + g.code_nx;
      } }
    CU_TRY(cudaMemsetAsync(h->d_overflow.p, 0, 16, s));
    CU_TRY(cudaMemcpyAsync(dl.p, &g, sizeof(g), cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(dc.p, cells.data(), sizeof(CellDesc) * nc, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(dcnt.p, counts.data(), 2 * nc, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(dslots.p, packed.data(), 4 * (size_t)ncand, cudaMemcpyHostToDevice, s));
    // form: ORBX_QT_FUSED = 1 / 0 forces the one-launch kernel of the latency form / the sort + tree pair; default = what a single frame gets (read per call: tests switch it)
    QfPlan q{}; q.threads = QF_THREADS; q.pool_cap = std::max(1056, align_up(tcap, 32)); q.cell_cap = nc; q.tab_cap = 0;
    const size_t qbudget = 224 * 1024, qfixed = qf_fixed_bytes(q.pool_cap, q.cell_cap, 0, q.threads);
    bool fused = q.pool_cap <= QF_MAXPOOL && qfixed + 2048 * 16 <= qbudget;
    { const char* e = std::getenv("ORBX_QT_FUSED"); if (e && std::atoi(e) == 0) fused = false; }
    if (fused) {
        q.key_cap = (int)std::min<size_t>(8192, ((qbudget - qfixed) / 16) & ~(size_t)31);
        { const char* e = std::getenv("ORBX_QT_KEYCAP"); if (e && std::atoi(e) >= 32) q.key_cap = std::min(q.key_cap, std::atoi(e) & ~31); }   // small values exercise the global-scratch path
        q.smem_bytes = (int)qf_smem_bytes(q);
        QfLevels ql{}; ql.lv[0] = g;
        k_octree_fused<QF_THREADS><<<dim3(1, 1), QF_THREADS, q.smem_bytes, s>>>(ql, dc.p, nc, ncand, ncand, g.kp_cap, 1, q, dslots.p, dcnt.p, doc.p, dsk.p, dspk.p, dn.p, dkp.p, dkc.p, h->d_overflow.p);
        LAUNCH_CHECK();
    } else {
    k_octree_sort_t<SORT_THREADS><<<dim3(1, 1), SORT_THREADS, octree_sort_smem_bytes(h->sort_smem_keys), s>>>(dl.p, dc.p, nc, ncand, ncand, 1, h->sort_smem_keys, dslots.p, dcnt.p, doc.p, dsk.p, dspk.p, dn.p);
    LAUNCH_CHECK();
    const int code_cap = 4096;
    const size_t tsm = (((size_t)tcap * (8 + 8 + 4 + 4 + 4 + 2 + 2 + 2 + 1) + 15) & ~(size_t)15) + (size_t)code_cap * 4 + 16;
    if (tcap <= PTREE_MAXCAP)
        k_octree_tree_par<<<dim3(1, 1), PTREE_THREADS, ptree_smem_bytes(tcap, code_cap), s>>>(dl.p, 1, ncand, g.kp_cap, tcap, code_cap, dsk.p, dspk.p, dn.p, dkp.p, dkc.p, h->d_overflow.p);
    else
        k_octree_tree<<<dim3(1, 1), 32, tsm, s>>>(dl.p, 1, ncand, g.kp_cap, tcap, code_cap, dsk.p, dspk.p, dn.p, dkp.p, dkc.p, h->d_overflow.p);
    LAUNCH_CHECK();
    }
    int n = 0, ovf = 0;
    std::vector<uint32_t> res(g.kp_cap);
    CU_TRY(cudaMemcpyAsync(&n, dkc.p, 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(&ovf, h->d_overflow.p, 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(res.data(), dkp.p, 4 * (size_t)g.kp_cap, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    if (ovf) { CU_TRY(cudaMemsetAsync(h->d_overflow.p, 0, 16, s)); FAIL(ORBX_E_OVERFLOW, "internal bound exceeded in the quadtree stage"); }
    *n_out = n;
    if (n > cap || !out) FAIL(ORBX_E_CAPACITY, "output buffer too small");
    for (int i = 0; i < n; ++i) {
        out[i].x = (float)(res[i] & 0xFFF); out[i].y = (float)((res[i] >> 12) & 0xFFF); out[i].size = 7.f; out[i].angle = -1.f;
        out[i].response = (float)(res[i] >> 24); out[i].octave = 0; out[i].class_id = -1;
    }
    return ORBX_OK;
}

extern "C" {

int orbx_debug_pyramid_level(orbx_extractor* h, int b, int level, uint8_t* dst, size_t dst_step) {
    if (!h || !h->have_pyramid || b < 0 || b >= h->lastB || level < 0 || level >= h->nlevels || !dst) FAIL(ORBX_E_INVALID, "bad arguments");
    CU_TRY(cudaSetDevice(h->device));
    const LevelGeom& g = h->levels[level];
    const uint8_t* base; int pitch;
    if (level == 0) { base = h->view.l0 + (long long)b * h->view.l0_fstride; pitch = h->view.l0_pitch; }
    else { base = h->d_pyr.p + (size_t)b * h->pyr_fstride + g.off; pitch = g.pitch; }
    return copy_level(h, base, pitch, level, 0, dst, dst_step);
}

int orbx_debug_blurred_level(orbx_extractor* h, int b, int level, uint8_t* dst, size_t dst_step) {
    if (!h || !h->have_pyramid || b < 0 || b >= h->lastB || level < 0 || level >= h->nlevels || !dst) FAIL(ORBX_E_INVALID, "bad arguments");
    CU_TRY(cudaSetDevice(h->device));
    int rc; if ((rc = run_blur(h, h->lastB))) return rc;
    const LevelGeom& g = h->levels[level];
    return copy_level(h, h->d_blur.p + (size_t)b * h->pyr_fstride + g.off, g.pitch, level, 0, dst, dst_step);
}

int orbx_debug_level_candidates(orbx_extractor* h, int b, int level, orbx_keypoint* out, int cap, int* n_out) {
    if (!h || !h->have_pyramid || b < 0 || b >= h->lastB || level < 0 || level >= h->nlevels || !n_out) FAIL(ORBX_E_INVALID, "bad arguments");
    CU_TRY(cudaSetDevice(h->device));
    const LevelGeom& g = h->levels[level];
    int n = 0;
    CU_TRY(cudaMemcpyAsync(&n, h->d_ncand.p + (size_t)b * h->nlevels + level, 4, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    *n_out = n;
    if (n <= 0) return ORBX_OK;
    if (n > cap || !out) FAIL(ORBX_E_CAPACITY, "candidate buffer too small");
    std::vector<uint32_t> p(n);
    CU_TRY(cudaMemcpyAsync(p.data(), h->d_ocand.p + (size_t)b * h->cand_per_frame + g.cand_off, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < n; ++i) {
        out[i].x = (float)(p[i] & 0xFFF); out[i].y = (float)((p[i] >> 12) & 0xFFF); out[i].size = 7.f; out[i].angle = -1.f;
        out[i].response = (float)(p[i] >> 24); out[i].octave = 0; out[i].class_id = -1;
    }
    return ORBX_OK;
}


__global__ void k_tap_sincos(unsigned lo, int n, float* __restrict__ s, float* __restrict__ c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) det_sincos(__uint_as_float(lo + (unsigned)i), s + i, c + i);
}
// det_sincos (the rBRIEF rotation's sin / cos, det_math.cuh) of the n floats whose bit patterns are lo_bits, lo_bits + 1, ...
int orbx_debug_sincos(orbx_extractor* h, unsigned lo_bits, int n, float* sin_out, float* cos_out) {
    if (!h || n <= 0 || !sin_out || !cos_out) FAIL(ORBX_E_INVALID, "bad arguments");
    CU_TRY(cudaSetDevice(h->device));
    DevBuf<float> ds, dc;
    struct Guard { DevBuf<float>&a, &b; ~Guard() { a.release(); b.release(); } } guard{ds, dc};
    if (ds.ensure(n) || dc.ensure(n)) return ORBX_E_CUDA;
    k_tap_sincos<<<(n + 255) / 256, 256, 0, h->stream>>>(lo_bits, n, ds.p, dc.p);
    LAUNCH_CHECK();
    CU_TRY(cudaMemcpyAsync(sin_out, ds.p, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaMemcpyAsync(cos_out, dc.p, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

}  // extern "C"

// Guard-zone check (ORBX_CANARY=1, see DevBuf in orbx_extractor.cu): returns the number of guarded device buffers whose 256-byte zones in
// front of or behind the payload no longer hold the pattern (0 = clean), or a negative status.  *n_blocks = buffers checked.
extern "C" int orbx_debug_canary_check(int* n_blocks) {
    if (n_blocks) *n_blocks = 0;
    if (!CanaryRegistry::on()) FAIL(ORBX_E_STATE, "ORBX_CANARY is not set: buffers carry no guard zones");
    CU_TRY(cudaDeviceSynchronize());
    std::vector<std::pair<uint8_t*, size_t>> blocks;
    { std::lock_guard<std::mutex> lk(CanaryRegistry::get().mu); blocks = CanaryRegistry::get().blocks; }
    int bad = 0;
    std::vector<uint8_t> g(2 * ORBX_GUARD);
    for (const auto& b : blocks) {
        CU_TRY(cudaMemcpy(g.data(), b.first, ORBX_GUARD, cudaMemcpyDeviceToHost));
        CU_TRY(cudaMemcpy(g.data() + ORBX_GUARD, b.first + ORBX_GUARD + b.second, ORBX_GUARD, cudaMemcpyDeviceToHost));
        bool ok = true;
        for (uint8_t v : g) ok = ok && v == 0xA5;
        bad += !ok;
    }
    if (n_blocks) *n_blocks = (int)blocks.size();
    return bad;
}
