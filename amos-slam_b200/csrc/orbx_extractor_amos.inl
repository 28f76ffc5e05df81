// orbx_extractor_amos.inl -- the Amos-SLAM stage of the extractor: orbx_cull (MovingKeyPoints) and the batched masked extraction
// (config C5); part of orbx_extractor.cu

static int upload_ellipse() {
    // cv::getStructuringElement(MORPH_ELLIPSE, 31x31): dx = cvRound(c * sqrt((r*r - dy*dy) * inv_r2))
    int dxs[31];
    const int r = 15, c = 15; const double inv_r2 = 1.0 / ((double)r * r);
    for (int i = 0; i < 31; ++i) {
        const int dy = i - r; dxs[i] = (int)lrint(c * std::sqrt((r * r - dy * dy) * inv_r2));
        if (dxs[i] != ell_dx(dy < 0 ? -dy : dy)) FAIL(ORBX_E_STATE, "ellipse table differs from the compiled-in one");
    }
    CU_TRY(cudaMemcpyToSymbol(c_ell_dx, dxs, sizeof(dxs)));
    return ORBX_OK;
}

// bit-pack the masks of frames [b0, b0+nframes) (u8, pitch multiple of 32; d_mask points at frame b0) and close them on h->cur;
// result: h->d_bits1 = dilate(mask != 0), [B][rows][wpr] (the erosion half of the closing is evaluated per keypoint, k_cull.cuh).  The caller has sized
// d_bits0 / d_bits1 for the whole batch.
static int run_closing(orbx_extractor* h, const uint8_t* d_mask, long long fstride, int pitch, int b0, int nframes, int rows, int cols, bool prepacked = false) {
    const int wpr = (cols + 31) / 32;
    cudaStream_t s = h->cur;
    uint32_t* a = h->d_bits0.p + (size_t)b0 * rows * wpr; uint32_t* b = h->d_bits1.p + (size_t)b0 * rows * wpr;
    dim3 grid((wpr + 63) / 64, rows, nframes);
    if (!prepacked) {                                             // (the host-pointer batch call uploads masks already packed, host_pack.cpp)
        k_mask_pack<<<grid, 64, 0, s>>>(d_mask, fstride, pitch, rows, cols, a, wpr);
        LAUNCH_CHECK();
    }
    k_bin_dilate31<<<dim3((wpr + DIL_TW - 1) / DIL_TW, (rows + DIL_TR - 1) / DIL_TR, nframes), 256, 0, s>>>(a, b, wpr, rows, cols);
    LAUNCH_CHECK();
    return ORBX_OK;
}
static int ensure_closing(orbx_extractor* h, int nframes, int rows, int cols) {
    int rc;
    if ((rc = upload_ellipse())) return rc;
    const size_t words = (size_t)nframes * rows * ((cols + 31) / 32);
    if (h->d_bits0.ensure(words) || h->d_bits1.ensure(words)) return ORBX_E_CUDA;
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
    const int pitch = align_up(cols, 128);
    if (n && !kp_inout) FAIL(ORBX_E_INVALID, "null keypoints");
    if (h->d_mask.ensure((size_t)pitch * rows + 64) || h->d_kp_tmp.ensure((size_t)std::max(n, 1) * 2) || h->d_desc_tmp.ensure((size_t)std::max(n, 1) * 32))
        return ORBX_E_CUDA;
    cudaStream_t s = h->stream;
    CU_TRY(cudaMemcpy2DAsync(h->d_mask.p, pitch, mask, mask_step, cols, rows, cudaMemcpyHostToDevice, s));
    // closing = erode(dilate(mask))   (:1697-1704): binary dilation of the bit-packed mask, erosion per keypoint (k_cull.cuh)
    if ((rc = ensure_closing(h, 1, rows, cols))) return rc;
    if ((rc = run_closing(h, h->d_mask.p, (long long)pitch * rows, pitch, 0, 1, rows, cols))) return rc;
    const int wpr = (cols + 31) / 32;
    if (n == 0) { CU_TRY(cudaStreamSynchronize(s)); return ORBX_OK; }
    // One pinned block up: keypoints, their scale (level 0 -> 1, else mvScaleFactor[level], :1712-1715) and the super-pixel term of the
    // test, rm_vector[centers[label(p) - 1].id] == 1 (:1722-1736).  That term reads three caller-side tables at N positions, so it is
    // evaluated while marshalling (N look-ups) instead of uploading the 8 B/pixel label map; p = (int)(pt * scale) is the same float
    // product the kernel forms for the mask term.  Out-of-range label / id indices (undefined behaviour in the reference) = not flagged.
    const size_t o_sc = (size_t)n * sizeof(KpOut), o_rm = o_sc + (size_t)n * 4, up_bytes = (o_rm + n + 15) & ~(size_t)15;
    if (h->h_gather_cap < up_bytes) {
        if (h->h_gather) cudaFreeHost(h->h_gather);
        h->h_gather = nullptr; h->h_gather_cap = 0;
        CU_TRY(cudaHostAlloc((void**)&h->h_gather, up_bytes + (up_bytes >> 1), cudaHostAllocDefault));
        h->h_gather_cap = up_bytes + (up_bytes >> 1);
    }
    if (h->d_gather.ensure(up_bytes)) return ORBX_E_CUDA;
    std::memcpy(h->h_gather, kp_inout, o_sc);
    float* hs = reinterpret_cast<float*>(h->h_gather + o_sc); uint8_t* hr = h->h_gather + o_rm;
    const size_t lpitch = label_step / sizeof(double);
    { int o = 0; for (int l = 0; l < h->nlevels; ++l) for (int i = 0; i < level_counts[l]; ++i, ++o) {
        const float sc = l ? h->mvScaleFactor[l] : 1.f;
        hs[o] = sc;
        volatile float fx = kp_inout[o].x * sc, fy = kp_inout[o].y * sc;          // one rounded float product each, as cv::Point2f * float
        const int px = (int)fx, py = (int)fy;
        uint8_t flag = 0;
        if (px >= 0 && py >= 0 && px < cols && py < rows) {
            const double idxd = label[(size_t)py * lpitch + px] - 1.0;
            if (idxd >= 0.0 && idxd < (double)ncenters) { const int id = centers_id[(size_t)idxd]; if (id >= 0 && id < nrm && rm_vector[id] == 1) flag = 1; }
        }
        hr[o] = flag;
    } }
    CU_TRY(cudaMemcpyAsync(h->d_gather.p, h->h_gather, up_bytes, cudaMemcpyHostToDevice, s));
    uint8_t* d_flags = h->d_desc_tmp.p;
    k_cull_flags<<<(n + 127) / 128, 128, 0, s>>>(reinterpret_cast<const KpIn*>(h->d_gather.p), reinterpret_cast<const float*>(h->d_gather.p + o_sc), n, h->d_bits1.p, wpr,
                                                  h->d_gather.p + o_rm, rows, cols, d_flags);
    LAUNCH_CHECK();
    // the flags land in the pinned block (the upload has been consumed by then: same stream)
    CU_TRY(cudaMemcpyAsync(h->h_gather, d_flags, n, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    const uint8_t* flags = h->h_gather;
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
// dynamic mask and, when lv.labels is set, the super-pixel term  ->  ProcessDesp.  Frames [b0, b0+nb) on h->cur; d_masks points at frame b0's mask.
static int run_masked_range(orbx_extractor* h, int b0, int nb, const uint8_t* d_masks, long long mfs, int mpitch, int rows, int cols, LabelView lv,
                            KpOut* d_kp, uint8_t* d_desc, int cap, int* d_counts, int* d_culled, bool prepacked, bool fork) {
    int rc;
    // fork (device-batch call, one stream): the closing of the masks needs nothing from the detection and the blur only the pyramid, so both go to the second
    // compute stream -- the closing beside resize / FAST, the blur beside the latency-bound quadtree -- and join in front of the culling.  (The host-pointer
    // pipeline already overlaps its chunks over four streams and does not fork.)
    fork = fork && !h->profiling && h->cur == h->stream;
    if (fork) {
        CU_TRY(cudaEventRecord(h->ev_fork, h->stream));
        CU_TRY(cudaStreamWaitEvent(h->s_alt, h->ev_fork, 0));
        h->cur = h->s_alt;
        rc = run_closing(h, d_masks, mfs, mpitch, b0, nb, rows, cols, prepacked);
        h->cur = h->stream;
        if (rc) return rc;
        if ((rc = run_detect(h, b0, nb, true))) return rc;                       // blur queued on s_alt behind the closing, after FAST (ev_fork); ev_join follows it
        CU_TRY(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    } else {
    if ((rc = run_detect(h, b0, nb))) return rc;
    if ((rc = run_closing(h, d_masks, mfs, mpitch, b0, nb, rows, cols, prepacked))) return rc;
    }
    if (d_culled) CU_TRY(cudaMemsetAsync(d_culled + b0, 0, (size_t)nb * sizeof(int), h->cur));
    const int wpr = (cols + 31) / 32;
    k_cull_levelkp<<<dim3(h->nlevels, nb), 32, 0, h->cur>>>(h->d_levels.p, h->nlevels, h->kp_per_frame, h->d_kp_level.p + (size_t)b0 * h->kp_per_frame,
                                                              h->d_kp_count.p + (size_t)b0 * h->nlevels, h->d_bits1.p + (size_t)b0 * rows * wpr, wpr, rows, cols,
                                                              lv, d_culled ? d_culled + b0 : nullptr);      // lv points at frame b0's labels / flags
    LAUNCH_CHECK();
    if (!fork && (rc = run_blur_range(h, b0, nb))) return rc;
    return run_orient(h, b0, nb, true, d_kp, d_desc, cap, d_counts, nullptr);
}

static int check_labels(const orbx_labels* L, int rows, int cols) {
    if (!L) return ORBX_OK;
    if (!L->labels || !L->flagged || L->n_labels <= 0 || L->n_labels > 65535 || L->label_step < (size_t)cols || L->label_frame_stride < L->label_step * (size_t)(rows - 1) + (size_t)cols)
        FAIL(ORBX_E_INVALID, "bad label arguments");
    return ORBX_OK;
}

extern "C" int orbx_extract_masked_batch_labels_device(orbx_extractor* h, const uint8_t* d_images, const uint8_t* d_masks, const orbx_labels* d_labels, int B, int rows, int cols,
                                                       size_t step, size_t frame_stride, size_t mask_step, size_t mask_frame_stride,
                                                       orbx_keypoint* d_kp_out, uint8_t* d_desc_out, int cap, int* d_counts_out, int* d_culled_out) {
    int rc = check_args(h, d_images, rows, cols, step); if (rc) return rc;
    if ((rc = check_labels(d_labels, rows, cols))) return rc;
    if (B <= 0 || !d_masks || !d_kp_out || !d_desc_out || !d_counts_out || cap <= 0 || mask_step < (size_t)cols) FAIL(ORBX_E_INVALID, "bad batch arguments");
    if ((rc = build_plan(h, rows, cols))) return rc;
    if ((rc = ensure_capacity(h, B, 0))) return rc;
    if ((rc = ensure_closing(h, B, rows, cols))) return rc;
    if (((uintptr_t)d_images & 15) == 0 && (step & 15) == 0 && (frame_stride & 15) == 0) {      // TMA reads level 0: 16-byte alignment
        h->view.l0 = d_images; h->view.l0_fstride = (long long)frame_stride; h->view.l0_pitch = (int)step;       // alias the caller's frames
    } else {
        const LevelGeom& g0 = h->levels[0];
        if ((rc = repitch_frames(h, h->d_pyr.p + g0.off, h->pyr_fstride, g0.pitch, d_images, (long long)frame_stride, (long long)step, cols, rows, B, h->stream))) return rc;
        h->view.l0 = h->d_pyr.p + g0.off; h->view.l0_fstride = h->pyr_fstride; h->view.l0_pitch = g0.pitch;
    }
    h->view.pyr = h->d_pyr.p; h->view.pyr_fstride = h->pyr_fstride;
    // masks: pack needs 32-byte row alignment; repack into our own pitched buffer unless the caller's layout already qualifies
    const uint8_t* mk = d_masks; long long mfs = (long long)mask_frame_stride; int mpitch = (int)mask_step;
    if (((uintptr_t)d_masks & 15) || (mask_step & 31) || (mask_frame_stride & 15) || mask_step < (size_t)align_up(cols, 32)) {
        mpitch = align_up(cols, 128); mfs = (long long)mpitch * rows;
        if (h->d_mask.ensure((size_t)mfs * B + 64)) return ORBX_E_CUDA;
        if ((rc = repitch_frames(h, h->d_mask.p, mfs, mpitch, d_masks, (long long)mask_frame_stride, (long long)mask_step, cols, rows, B, h->stream))) return rc;
        mk = h->d_mask.p;
    }
    h->cur = h->stream;
    LabelView lv{nullptr, 0, 0, nullptr, 0};
    if (d_labels) lv = LabelView{d_labels->labels, (long long)d_labels->label_frame_stride, (int)d_labels->label_step, d_labels->flagged, d_labels->n_labels};
    rc = run_masked_range(h, 0, B, mk, mfs, mpitch, rows, cols, lv, reinterpret_cast<KpOut*>(d_kp_out), d_desc_out, cap, d_counts_out, d_culled_out, false, true);
    if (!rc) { h->lastB = B; h->blur_valid = true; }
    return rc;
}
extern "C" int orbx_extract_masked_batch_device(orbx_extractor* h, const uint8_t* d_images, const uint8_t* d_masks, int B, int rows, int cols,
                                                size_t step, size_t frame_stride, size_t mask_step, size_t mask_frame_stride,
                                                orbx_keypoint* d_kp_out, uint8_t* d_desc_out, int cap, int* d_counts_out, int* d_culled_out) {
    return orbx_extract_masked_batch_labels_device(h, d_images, d_masks, nullptr, B, rows, cols, step, frame_stride, mask_step, mask_frame_stride, d_kp_out, d_desc_out, cap, d_counts_out, d_culled_out);
}

// host-pointer form: same chunked H2D -> compute -> D2H pipeline as orbx_extract_batch, with the masks riding along
extern "C" int orbx_extract_masked_batch(orbx_extractor* h, const uint8_t* images, const uint8_t* masks, int B, int rows, int cols,
                                         size_t step, size_t frame_stride, size_t mask_step, size_t mask_frame_stride,
                                         orbx_keypoint* kp_out, uint8_t* desc_out, int cap, int* counts_out, int* culled_out) {
    if (!masks || mask_step < (size_t)cols) FAIL(ORBX_E_INVALID, "bad mask arguments");
    return host_batch_pipeline(h, images, masks, nullptr, B, rows, cols, step, frame_stride, mask_step, mask_frame_stride, kp_out, desc_out, cap, counts_out, culled_out);
}
extern "C" int orbx_extract_masked_batch_labels(orbx_extractor* h, const uint8_t* images, const uint8_t* masks, const orbx_labels* labels, int B, int rows, int cols,
                                                size_t step, size_t frame_stride, size_t mask_step, size_t mask_frame_stride,
                                                orbx_keypoint* kp_out, uint8_t* desc_out, int cap, int* counts_out, int* culled_out) {
    if (!masks || mask_step < (size_t)cols) FAIL(ORBX_E_INVALID, "bad mask arguments");
    int rc = check_labels(labels, rows, cols); if (rc) return rc;
    return host_batch_pipeline(h, images, masks, labels, B, rows, cols, step, frame_stride, mask_step, mask_frame_stride, kp_out, desc_out, cap, counts_out, culled_out);
}

