// Host driver of the linear-space aligners (included by seqa_cuda.cu): level-synchronous recursion.
namespace {

template <class T> int ls_alloc(T **p, uint64_t n)
{
    if (*p) cudaFree(*p);
    *p = nullptr;
    if (cudaMalloc((void **)p, std::max<uint64_t>(n, 1) * sizeof(T)) != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(SEQA_ERR_NOMEM, "device allocation of %llu bytes failed", (unsigned long long)(n * sizeof(T)));
    }
    return SEQA_OK;
}

void ls_release(LsState &ls)
{
    if (ls.d_nodes[0]) cudaFree(ls.d_nodes[0]);
    if (ls.d_nodes[1]) cudaFree(ls.d_nodes[1]);
    if (ls.d_count) cudaFree(ls.d_count);
    if (ls.d_overflow) cudaFree(ls.d_overflow);
    if (ls.d_sweeps) cudaFree(ls.d_sweeps);
    if (ls.d_tasks) cudaFree(ls.d_tasks);
    if (ls.d_prog) cudaFree(ls.d_prog);
    if (ls.d_rows) cudaFree(ls.d_rows);
    if (ls.d_row_off) cudaFree(ls.d_row_off);
    if (ls.d_row_w) cudaFree(ls.d_row_w);
    if (ls.d_idx) cudaFree(ls.d_idx);
    ls = LsState();
}

int ls_plan(seqa_ctx *c, const std::vector<uint32_t> &len1, const std::vector<uint32_t> &len2,
            const std::vector<uint32_t> &idx, bool myers_miller)
{
    LsState &ls = c->ls;
    const seqa_params &prm = c->prm;
    const int sms = c->sms;
    const uint64_t n = idx.size();
    ls.mm = myers_miller;
    ls.roots.resize(n);
    ls.row_off.assign(len1.size(), 0);
    ls.row_w.assign(len1.size(), 0);
    const uint64_t narr = myers_miller ? LsArr<true>::COUNT : LsArr<false>::COUNT;
    const int forced = (prm.flags & SEQA_FLAG_LS_R1) ? 1 : 0;
    const uint64_t rpb_min = 32; // smallest row block any sweep may be cut into
    // every value of a sweep stays inside (-2^29, 2^29): lets the kernels use one large negative constant for the
    // reference's INT_MIN diagonal candidate (include/SANeedlemanWunsch.h:138) without overflow
    const int64_t unit = (int64_t)std::max(std::abs((int64_t)prm.gap), std::abs((int64_t)prm.gap_open)) +
                         std::abs((int64_t)prm.gap_extend) + std::abs((int64_t)prm.match) +
                         (prm.allow_mismatch ? std::abs((int64_t)prm.mismatch) : 0);
    uint64_t run = 0, sum_m = 0, blocks = 0;
    for (uint64_t k = 0; k < n; k++) {
        const uint32_t p = idx[k];
        const uint64_t M = len1[p], N = len2[p];
        if (M + N > 0x3fffffffull) return fail(SEQA_ERR_UNSUPPORTED, "pair %u is too long for 32-bit cell indices", p);
        if ((int64_t)(M + N + 4) * unit >= ((int64_t)1 << 29))
            return fail(SEQA_ERR_UNSUPPORTED, "pair %u: (len1+len2) x scoring magnitude exceeds the 2^29 score range", p);
        // an internal node at depth d sits at column offset j0 + q, q < 2^d <= M: one array needs N + M + 2 ints
        const uint64_t w = N + M + 2;
        ls.row_off[p] = run;
        ls.row_w[p] = (uint32_t)w;
        run += narr * w;
        sum_m += M;
        blocks += M / rpb_min + 2;
        LsNode r;
        r.pair = (int)p;
        r.i0 = 0; r.m = (int)M; r.j0 = 0; r.n = (int)N; r.q = 0;
        r.tb = r.te = 0; // Myers-Miller entry tb = te = GapOpen is filled in at run time (include/SAMyersMiller.h:417)
        ls.roots[k] = r;
    }
    (void)forced;
    ls.rows_total = run;
    ls.node_cap = sum_m + n + 64;
    ls.sum_blocks = blocks;
    // tasks of one level: sum over nodes of ceil(rows_f / rpb) + ceil(rows_r / rpb) <= sum(m) / rpb + 2 * nodes
    ls.task_cap = std::min<uint64_t>(blocks + 2 * ls.node_cap + 8, 0xfffffff0ull);
    if (ls.node_cap > ls.cap_nodes) {
        CKS(ls_alloc(&ls.d_nodes[0], ls.node_cap));
        CKS(ls_alloc(&ls.d_nodes[1], ls.node_cap));
        CKS(ls_alloc(&ls.d_sweeps, 2 * ls.node_cap));
        ls.cap_nodes = ls.node_cap;
    }
    if (ls.task_cap > ls.cap_tasks) {
        CKS(ls_alloc(&ls.d_tasks, ls.task_cap));
        CKS(ls_alloc(&ls.d_prog, ls.task_cap));
        ls.cap_tasks = ls.task_cap;
    }
    if (ls.rows_total > ls.cap_rows) {
        CKS(ls_alloc(&ls.d_rows, ls.rows_total));
        ls.cap_rows = ls.rows_total;
    }
    if (len1.size() > ls.cap_pairs) {
        CKS(ls_alloc(&ls.d_row_off, len1.size()));
        CKS(ls_alloc(&ls.d_row_w, len1.size()));
        CKS(ls_alloc(&ls.d_idx, len1.size()));
        ls.cap_pairs = len1.size();
    }
    if (!ls.d_count) CKS(ls_alloc(&ls.d_count, 4));
    if (!ls.d_overflow) CKS(ls_alloc(&ls.d_overflow, 1));
    if (!ls.sweep_blocks) {
#ifdef SEQA_EMU
        ls.sweep_blocks = ls.sweep2_blocks = ls.sweep2_blocks_hb = sms;
#else
        int nb_hb = 0, nb_mm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb_hb, ls_sweep_kernel<false>, LS_BLOCK, 0));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb_mm, ls_sweep_kernel<true>, LS_BLOCK, 0));
        ls.sweep_blocks = sms * std::max(1, std::min(nb_hb, nb_mm));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb_hb, ls_sweep2_kernel<false>, LS_BLOCK, 0));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb_mm, ls_sweep2_kernel<true>, LS_BLOCK, 0));
        // per variant: Hirschberg's 72-80 registers leave room for more resident warps than MyersMiller's 96-116
        // (round 1 launched both with the smaller of the two: 16 warps per SM, 23 % of the slots)
        const int bps = env_int("SEQA_LS_BPS", 0, 0, 16);
        ls.sweep2_blocks_hb = sms * std::max(1, bps ? bps : nb_hb);
        ls.sweep2_blocks = sms * std::max(1, bps ? bps : nb_mm);
#endif
    }
    if (n) {
        // staged in pinned memory and copied on the ctx's upload stream (ordered before the kernels), like the packed
        // plan's perm / jobs: no legacy-stream copy that would synchronise with the caller's other streams
        CKS(c->ls_rowoff_pin.ensure(len1.size()));
        CKS(c->ls_roww_pin.ensure(len1.size()));
        CKS(c->ls_idx_pin.ensure(n));
        std::copy(ls.row_off.begin(), ls.row_off.end(), c->ls_rowoff_pin.p);
        std::copy(ls.row_w.begin(), ls.row_w.end(), c->ls_roww_pin.p);
        std::copy(idx.begin(), idx.end(), c->ls_idx_pin.p);
        CK(cudaMemcpyAsync(ls.d_row_off, c->ls_rowoff_pin.p, len1.size() * 8, cudaMemcpyHostToDevice, c->up));
        CK(cudaMemcpyAsync(ls.d_row_w, c->ls_roww_pin.p, len1.size() * 4, cudaMemcpyHostToDevice, c->up));
        CK(cudaMemcpyAsync(ls.d_idx, c->ls_idx_pin.p, n * 4, cudaMemcpyHostToDevice, c->up));
        CKS(order_after(c, c->up, c->stream));
    }
    return SEQA_OK;
}

int ls_run(seqa_ctx *c, bool want_ops)
{
    (void)want_ops;
    LsState &ls = c->ls;
    const uint64_t n = c->lidx.size();
    if (n == 0) return SEQA_OK;
    for (auto &r : ls.roots) r.tb = r.te = c->prm.gap_open;
    CK(cudaMemsetAsync(c->slots.p, LS_HOLE, c->slots_total, c->stream));
    CK(cudaMemsetAsync(ls.d_overflow, 0, sizeof(int), c->stream));
    CK(cudaMemcpyAsync(ls.d_nodes[0], ls.roots.data(), n * sizeof(LsNode), cudaMemcpyHostToDevice, c->stream));
    LsArgs A{};
    A.bases = c->bases.p;
    A.off1 = c->off1.p;
    A.off2 = c->off2.p;
    A.len1 = c->len1.p;
    A.len2 = c->len2.p;
    A.cnt = ls.d_count;
    A.out_cap = (uint32_t)std::min<uint64_t>(ls.node_cap, 0xffffffffull);
    A.overflow = ls.d_overflow;
    A.sweeps = ls.d_sweeps;
    A.tasks = ls.d_tasks;
    A.task_cap = (uint32_t)ls.task_cap;
    A.prog = ls.d_prog;
    A.rows = ls.d_rows;
    A.row_off = ls.d_row_off;
    A.row_w = ls.d_row_w;
    A.slots = c->slots.p;
    A.slot_off = c->slot_off.p;
    A.sc = c->sc;
    A.force_r = (c->prm.flags & SEQA_FLAG_LS_R1) ? 1 : 0;
    // packed dual sweeps: small scoring constants (int8 profiles, 16-bit windows) and symbols in {A,C,G,T}
    {
        const seqa_params &q = c->prm;
        const int g = ls.mm ? -(q.gap_open + q.gap_extend) : -q.gap, x = q.allow_mismatch ? -q.mismatch : 0;
        bool ok = !(q.flags & SEQA_FLAG_FORCE_GENERIC) && q.match <= 100 && x <= 100 && g <= 50 && q.match + g <= 120 && x - g >= -120;
        if (ok) {
            CK(cudaMemsetAsync(c->flags.p + 2, 0, sizeof(int), c->stream));
            const unsigned blocks = (unsigned)std::min<uint64_t>((c->bases_len + 256 * 16 - 1) / (256 * 16) + 1, (uint64_t)c->sms * 8);
            LAUNCH(c, (ls_check_acgt_kernel), blocks, 256, 0, c->bases.p, c->bases_len, c->flags.p + 2);
            int bad = 0;
            CK(cudaMemcpyAsync(&bad, c->flags.p + 2, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            ok = !bad;
        }
        A.packed = ok ? 1 : 0;
    }
    uint32_t count = (uint32_t)n;
    int cur = 0;
    ls.levels_run = 0;
    cudaEventRecord(next_event(c), c->stream);
    while (count > 0) {
        A.in = ls.d_nodes[cur];
        A.n_in = count;
        A.out = ls.d_nodes[cur ^ 1];
        CK(cudaMemsetAsync(ls.d_count, 0, 4 * sizeof(uint32_t), c->stream));
        const uint64_t prog_n = std::min<uint64_t>(ls.sum_blocks + 2ull * count + 8, ls.task_cap);
        CK(cudaMemsetAsync(ls.d_prog, 0, prog_n * sizeof(int), c->stream));
        const unsigned egrid = std::min<unsigned>((count + 3) / 4, (unsigned)c->sms * 16);
        if (ls.mm) {
            LAUNCH(c, (ls_expand_kernel<true>), egrid, 128, 0, A);
            if (A.packed)
                LAUNCH(c, (ls_sweep2_kernel<true>), (unsigned)ls.sweep2_blocks, LS_BLOCK, 0, A);
            else
                LAUNCH(c, (ls_sweep_kernel<true>), (unsigned)ls.sweep_blocks, LS_BLOCK, 0, A);
            LAUNCH(c, (ls_split_kernel<true>), egrid, 128, 0, A);
        } else {
            LAUNCH(c, (ls_expand_kernel<false>), egrid, 128, 0, A);
            if (A.packed)
                LAUNCH(c, (ls_sweep2_kernel<false>), (unsigned)ls.sweep2_blocks_hb, LS_BLOCK, 0, A);
            else
                LAUNCH(c, (ls_sweep_kernel<false>), (unsigned)ls.sweep_blocks, LS_BLOCK, 0, A);
            LAUNCH(c, (ls_split_kernel<false>), egrid, 128, 0, A);
        }
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&count, ls.d_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        cur ^= 1;
        if (++ls.levels_run > 200) return fail(SEQA_ERR_CUDA, "internal: recursion does not terminate");
    }
    cudaEventRecord(next_event(c), c->stream);
    int ovf = 0;
    CK(cudaMemcpyAsync(&ovf, ls.d_overflow, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (ovf) return fail(SEQA_ERR_CUDA, "internal: node / task list overflow");
    LsFinishArgs F{};
    F.bases = c->bases.p;
    F.off1 = c->off1.p;
    F.off2 = c->off2.p;
    F.len1 = c->len1.p;
    F.len2 = c->len2.p;
    F.idx = ls.d_idx;
    F.count = n;
    F.slots = c->slots.p;
    F.slot_off = c->slot_off.p;
    F.slot_start = c->slot_start.p;
    F.ops_len = c->ops_len.p;
    F.start_i = c->start_i.p;
    F.start_j = c->start_j.p;
    F.end_i = c->end_i.p;
    F.end_j = c->end_j.p;
    F.score = c->score.p;
    F.sc = c->sc;
    F.affine = ls.mm ? 1 : 0;
    const unsigned blocks = (unsigned)std::min<uint64_t>((n + 3) / 4, (uint64_t)c->sms * 16);
    LAUNCH(c, (ls_finish_kernel), blocks, 128, 0, F);
    CK(cudaGetLastError());
    c->last_kernel = A.packed ? (ls.mm ? "ls_sweep2_mm_s16x2" : "ls_sweep2_hb_s16x2") : (ls.mm ? "ls_sweep_mm_i32" : "ls_sweep_hb_i32");
    return SEQA_OK;
}

} // namespace
