// orbx_extractor_debug.inl -- orbx_cull (MovingKeyPoints), the batched masked extraction (config C5) and the quadtree stage tap;
// part of orbx_extractor.cu

static int upload_ellipse() {
    // cv::getStructuringElement(MORPH_ELLIPSE, 31x31): dx = cvRound(c * sqrt((r*r - dy*dy) * inv_r2))
    int dxs[31];
    const int r = 15, c = 15; const double inv_r2 = 1.0 / ((double)r * r);
    for (int i = 0; i < 31; ++i) { const int dy = i - r; dxs[i] = (int)lrint(c * std::sqrt((r * r - dy * dy) * inv_r2)); }
    CU_TRY(cudaMemcpyToSymbol(c_ell_dx, dxs, sizeof(dxs)));
    return ORBX_OK;
}

// bit-pack nframes masks (u8, pitch multiple of 32) and close them; result: h->d_bits1 = (closing != 0), [nframes][rows][wpr]
static int run_closing(orbx_extractor* h, const uint8_t* d_mask, long long fstride, int pitch, int nframes, int rows, int cols, cudaStream_t s) {
    int rc;
    if ((rc = upload_ellipse())) return rc;
    const int wpr = (cols + 31) / 32;
    if (h->d_bits0.ensure((size_t)nframes * rows * wpr) || h->d_bits1.ensure((size_t)nframes * rows * wpr)) return ORBX_E_CUDA;
    dim3 grid((wpr + 63) / 64, rows, nframes);
    k_mask_pack<<<grid, 64, 0, s>>>(d_mask, fstride, pitch, rows, cols, h->d_bits0.p, wpr);
    LAUNCH_CHECK();
    k_bin_dilate31<false><<<grid, 64, 0, s>>>(h->d_bits0.p, h->d_bits1.p, wpr, rows, cols);
    LAUNCH_CHECK();
    k_bin_dilate31<true><<<grid, 64, 0, s>>>(h->d_bits1.p, h->d_bits0.p, wpr, rows, cols);
    LAUNCH_CHECK();
    std::swap(h->d_bits0, h->d_bits1);
    return ORBX_OK;
}

extern "C" int orbx_cull(orbx_extractor* h, const uint8_t* mask, size_t mask_step, const double* label, size_t label_step,
                         int rows, int cols, const int* centers_id, int ncenters, const int* rm_vector, int nrm,
                         orbx_keypoint* kp_inout, int* level_counts, orbx_keypoint* culled_out, int* n_culled) {
    if (n_culled) *n_culled = 0;
    if (!h || !mask || !label || !level_counts || !n_culled || rows <= 0 || cols <= 0 || ncenters < 0 || nrm < 0) FAIL(ORBX_E_INVALID, "bad arguments");
    if (mask_step < (size_t)cols || label_step < (size_t)cols * sizeof(double) || (label_step % sizeof(double))) FAIL(ORBX_E_INVALID, "bad steps");
    if ((ncenters && !centers_id) || (nrm && !rm_vector)) FAIL(ORBX_E_INVALID, "null table");
    CU_TRY(cudaSetDevice(h->device));
    int n = 0;
    for (int l = 0; l < h->nlevels; ++l) { if (level_counts[l] < 0) FAIL(ORBX_E_INVALID, "negative level count"); n += level_counts[l]; }
    int rc;
    if ((rc = upload_ellipse())) return rc;
    const int pitch = align_up(cols, 128);
    if (h->d_mask.ensure((size_t)pitch * rows + 64) || h->d_label.ensure((size_t)rows * cols) ||
        h->d_ids.ensure((size_t)ncenters + nrm + 4) || h->d_kp_tmp.ensure((size_t)std::max(n, 1) * 2) || h->d_desc_tmp.ensure((size_t)std::max(n, 1) * 32))
        return ORBX_E_CUDA;
    cudaStream_t s = h->stream;
    CU_TRY(cudaMemcpy2DAsync(h->d_mask.p, pitch, mask, mask_step, cols, rows, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpy2DAsync(h->d_label.p, (size_t)cols * 8, label, label_step, (size_t)cols * 8, rows, cudaMemcpyHostToDevice, s));
    if (ncenters) CU_TRY(cudaMemcpyAsync(h->d_ids.p, centers_id, (size_t)ncenters * 4, cudaMemcpyHostToDevice, s));
    if (nrm) CU_TRY(cudaMemcpyAsync(h->d_ids.p + ncenters, rm_vector, (size_t)nrm * 4, cudaMemcpyHostToDevice, s));
    // closing = erode(dilate(mask))   (:1697-1704), as two binary dilations on the bit-packed mask (k_cull.cuh)
    if ((rc = run_closing(h, h->d_mask.p, (long long)pitch * rows, pitch, 1, rows, cols, s))) return rc;
    const int wpr = (cols + 31) / 32;
    if (n == 0) { CU_TRY(cudaStreamSynchronize(s)); return ORBX_OK; }
    if (!kp_inout) FAIL(ORBX_E_INVALID, "null keypoints");
    // per-keypoint scale: level 0 -> 1, else mvScaleFactor[level]  (:1712-1715); scales ride in the desc scratch
    std::vector<float> scales(n);
    { int o = 0; for (int l = 0; l < h->nlevels; ++l) for (int i = 0; i < level_counts[l]; ++i) scales[o++] = l ? h->mvScaleFactor[l] : 1.f; }
    float* d_scales = reinterpret_cast<float*>(h->d_desc_tmp.p);
    uint8_t* d_flags = h->d_desc_tmp.p + (size_t)n * 4;
    CU_TRY(cudaMemcpyAsync(h->d_kp_tmp.p, kp_inout, sizeof(KpOut) * n, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(d_scales, scales.data(), (size_t)n * 4, cudaMemcpyHostToDevice, s));
    k_cull_flags<<<(n + 127) / 128, 128, 0, s>>>(reinterpret_cast<const KpIn*>(h->d_kp_tmp.p), d_scales, n, h->d_bits1.p, wpr, h->d_label.p, cols,
                                                  rows, cols, h->d_ids.p, ncenters, h->d_ids.p + ncenters, nrm, d_flags);
    LAUNCH_CHECK();
    std::vector<uint8_t> flags(n);
    CU_TRY(cudaMemcpyAsync(flags.data(), d_flags, n, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    // stable erase per level, culled keypoints appended in visiting order (marshalling of the device flags)
    int o = 0, w = 0, nc = 0;
    for (int l = 0; l < h->nlevels; ++l) {
        int kept = 0;
        for (int i = 0; i < level_counts[l]; ++i, ++o) {
            if (flags[o]) { if (culled_out) culled_out[nc] = kp_inout[o]; ++nc; }
            else { kp_inout[w++] = kp_inout[o]; ++kept; }
        }
        level_counts[l] = kept;
    }
    *n_culled = nc;
    return ORBX_OK;
}

// Batched Amos path (BASELINE config 5): per frame  operator()(img, mask, vector<vector<KeyPoint>>&)  ->  MovingKeyPoints with the
// dynamic mask (no super-pixel labels)  ->  ProcessDesp.  Device pointers, asynchronous on orbx_stream(h).
extern "C" int orbx_extract_masked_batch_device(orbx_extractor* h, const uint8_t* d_images, const uint8_t* d_masks, int B, int rows, int cols,
                                                size_t step, size_t frame_stride, size_t mask_step, size_t mask_frame_stride,
                                                orbx_keypoint* d_kp_out, uint8_t* d_desc_out, int cap, int* d_counts_out, int* d_culled_out) {
    int rc = check_args(h, d_images, rows, cols, step); if (rc) return rc;
    if (B <= 0 || !d_masks || !d_kp_out || !d_desc_out || !d_counts_out || cap <= 0 || mask_step < (size_t)cols) FAIL(ORBX_E_INVALID, "bad batch arguments");
    if ((rc = build_plan(h, rows, cols))) return rc;
    if ((rc = ensure_capacity(h, B, 0))) return rc;
    if (((uintptr_t)d_images & 3) || (step & 3) || (frame_stride & 3)) FAIL(ORBX_E_INVALID, "device frames must be 4-byte aligned (pointer, step, frame stride)");
    h->view.l0 = d_images; h->view.l0_fstride = (long long)frame_stride; h->view.l0_pitch = (int)step;
    h->view.pyr = h->d_pyr.p; h->view.pyr_fstride = h->pyr_fstride;
    cudaStream_t s = h->stream;
    // masks: pack needs 32-byte row alignment; repack into our own pitched buffer unless the caller's layout already qualifies
    const uint8_t* mk = d_masks; long long mfs = (long long)mask_frame_stride; int mpitch = (int)mask_step;
    if (((uintptr_t)d_masks & 15) || (mask_step & 31) || (mask_frame_stride & 15) || mask_step < (size_t)align_up(cols, 32)) {
        mpitch = align_up(cols, 128); mfs = (long long)mpitch * rows;
        if (h->d_mask.ensure((size_t)mfs * B + 64)) return ORBX_E_CUDA;
        for (int b = 0; b < B; ++b)
            CU_TRY(cudaMemcpy2DAsync(h->d_mask.p + (size_t)b * mfs, mpitch, d_masks + (size_t)b * mask_frame_stride, mask_step, cols, rows, cudaMemcpyDeviceToDevice, s));
        mk = h->d_mask.p;
    }
    if ((rc = run_detect(h, 0, B, !h->profiling))) return rc;              // blur forked next to FAST / quadtree / mask closing
    if ((rc = run_closing(h, mk, mfs, mpitch, B, rows, cols, s))) return rc;
    if (d_culled_out) CU_TRY(cudaMemsetAsync(d_culled_out, 0, (size_t)B * sizeof(int), s));
    k_cull_levelkp<<<dim3(h->nlevels, B), 32, 0, s>>>(h->d_levels.p, h->nlevels, h->kp_per_frame, h->d_kp_level.p, h->d_kp_count.p,
                                                       h->d_bits1.p, (cols + 31) / 32, rows, cols, d_culled_out);
    LAUNCH_CHECK();
    if (h->profiling) { if ((rc = run_blur(h, B))) return rc; }
    else { CU_TRY(cudaStreamWaitEvent(h->stream, h->ev_join, 0)); h->blur_valid = true; }
    return run_orient(h, 0, B, true, reinterpret_cast<KpOut*>(d_kp_out), d_desc_out, cap, d_counts_out, nullptr);
}

// host-pointer form: synchronous; frames and masks are uploaded, results downloaded
extern "C" int orbx_extract_masked_batch(orbx_extractor* h, const uint8_t* images, const uint8_t* masks, int B, int rows, int cols,
                                         size_t step, size_t frame_stride, size_t mask_step, size_t mask_frame_stride,
                                         orbx_keypoint* kp_out, uint8_t* desc_out, int cap, int* counts_out, int* culled_out) {
    int rc = check_args(h, images, rows, cols, step); if (rc) return rc;
    if (B <= 0 || !masks || !kp_out || !desc_out || !counts_out || cap <= 0 || mask_step < (size_t)cols) FAIL(ORBX_E_INVALID, "bad batch arguments");
    if ((rc = build_plan(h, rows, cols))) return rc;
    if ((rc = ensure_capacity(h, B, cap))) return rc;
    const int p0 = align_up(cols, 128);
    const size_t fs = (size_t)p0 * rows;
    if (h->d_l0.ensure(fs * B + 64) || h->d_mask.ensure(fs * B + 64) || h->d_culled.ensure(B)) return ORBX_E_CUDA;
    cudaStream_t s = h->stream;
    for (int b = 0; b < B; ++b) {
        CU_TRY(cudaMemcpy2DAsync(h->d_l0.p + b * fs, p0, images + (size_t)b * frame_stride, step, cols, rows, cudaMemcpyHostToDevice, s));
        CU_TRY(cudaMemcpy2DAsync(h->d_mask.p + b * fs, p0, masks + (size_t)b * mask_frame_stride, mask_step, cols, rows, cudaMemcpyHostToDevice, s));
    }
    if ((rc = orbx_extract_masked_batch_device(h, h->d_l0.p, h->d_mask.p, B, rows, cols, p0, fs, p0, fs, reinterpret_cast<orbx_keypoint*>(h->d_kp_out.p),
                                               h->d_desc_out.p, cap, h->d_counts.p, h->d_culled.p))) return rc;
    CU_TRY(cudaMemcpyAsync(kp_out, h->d_kp_out.p, (size_t)B * cap * sizeof(KpOut), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(desc_out, h->d_desc_out.p, (size_t)B * cap * 32, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(counts_out, h->d_counts.p, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (culled_out) CU_TRY(cudaMemcpyAsync(culled_out, h->d_culled.p, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, s));
    int ovf = 0;
    CU_TRY(cudaMemcpyAsync(&ovf, h->d_overflow.p, 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    if (ovf) FAIL(ORBX_E_OVERFLOW, "internal bound exceeded in the quadtree stage");
    return ORBX_OK;
}

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
    CU_TRY(cudaMemsetAsync(h->d_overflow.p, 0, 16, s));
    CU_TRY(cudaMemcpyAsync(dl.p, &g, sizeof(g), cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(dc.p, cells.data(), sizeof(CellDesc) * nc, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(dcnt.p, counts.data(), 2 * nc, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(dslots.p, packed.data(), 4 * (size_t)ncand, cudaMemcpyHostToDevice, s));
    k_octree_sort<<<dim3(1, 1), SORT_THREADS, octree_sort_smem_bytes(h->sort_smem_keys), s>>>(dl.p, dc.p, nc, ncand, ncand, 1, h->sort_smem_keys, dslots.p, dcnt.p, doc.p, dsk.p, dspk.p, dn.p);
    LAUNCH_CHECK();
    const int code_cap = 4096;
    const size_t tsm = (((size_t)tcap * (8 + 8 + 4 + 4 + 4 + 2 + 2 + 2 + 1) + 15) & ~(size_t)15) + (size_t)code_cap * 4 + 16;
    k_octree_tree<<<dim3(1, 1), 32, tsm, s>>>(dl.p, 1, ncand, g.kp_cap, tcap, code_cap, dsk.p, dspk.p, dn.p, dkp.p, dkc.p, h->d_overflow.p);
    LAUNCH_CHECK();
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
