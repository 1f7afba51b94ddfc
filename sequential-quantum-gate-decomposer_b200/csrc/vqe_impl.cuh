// vqe_impl.cuh -- host orchestration of the state-vector (VQE) path; included at the end of sqgpu.cu.
// Variational_Quantum_Eigensolver_Base::optimization_problem (…Base.cpp:1088-1121) and
// optimization_problem_combined_non_static (:1131-1199), batched over parameter sets.
#pragma once

// Windowed executor: the state makes one HBM round trip per SEGMENT (build_window_plan) instead of one per gate. Forward:
// fused_exec<MODE_APPLY> per segment, in place; energy as in the streaming path; gradient: fused_exec<MODE_BWD> per segment
// in reverse order on (psi, lambda), W' partials per (parameter set, CTA) reduced by reduce_partials.
// Returns 1 when the shared-memory plan does not fit (caller falls back to the streaming path).
static int vqe_window_dev(sqgpu_ctx* c, const double* d_params, int batch, bool with_grad, double* d_energy, double* d_grad, cudaStream_t st) {
    PlanScope keep(c);
    c->P = &c->planW;
    const int rows = c->rows, w = c->win_w, wr = 1 << w, wc = rows >> w;
    int rc;
    const int slice = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::min(batch, 65535), ((size_t)1 << 30) / ((size_t)rows * sizeof(cplx))));
    // four quarter-width CTAs per SM: with ~5 ops between a tile's load and its store, more independent CTAs in different
    // phases are what overlaps the HBM round trips with the tensor work (measured: +12 % energy, +7 % gradient over two)
    const FusedPlan pf = plan_fused(c, MODE_APPLY, wr, wc, std::min(slice, batch), 4, true);
    // backward segments: two 256-thread CTAs per SM with single-column tiles (a + beta: 64 KB each). With the tiles loaded by
    // asynchronous copies and dealt round robin (neighbouring tiles in flight together), two CTAs in different phases beat ONE
    // 512-thread CTA with a two-column tile: 46.7 against 51.7 ms per 64 sets on C5 (profiles/r2_variants_win2.jsonl; round 1
    // had it the other way round with register-staged loads)
    const FusedPlan pb = with_grad ? plan_fused(c, MODE_BWD, wr, wc, std::min(slice, batch), 2) : pf;
    if (!pf.ok || !pb.ok) return 1;
    const int nblk = std::min(c->sm_count * 8, std::max(1, rows / 512));
    if ((rc = c->wMat.ensure((size_t)2 * slice * rows * sizeof(cplx)))) return rc;
    if ((rc = c->wTrPart.ensure((size_t)slice * std::max(nblk * 32 + 6, pb.chunks * 6) * sizeof(double)))) return rc;
    if (with_grad) {
        if ((rc = c->wWPart.ensure(std::max<size_t>(1, (size_t)slice * pb.w_slices * c->P->w_total) * sizeof(cplx)))) return rc;
        if ((rc = c->wTraces.ensure((size_t)slice * (1 + c->n_params) * 6 * sizeof(double)))) return rc;
    }
    cplx* psi = c->wMat.as<cplx>();
    cplx* lam = psi + (size_t)slice * rows;
    auto seg_args = [&](const FusedPlan& p, const sqgpu_ctx::Segment& sg, ExecArgs& a) {
        fill_common_args(c, p, a, wr, wc);
        a.n = w;
        a.in = psi;
        a.out = psi;
        a.in_ystride = a.out_ystride = rows;
        a.wmask = sg.wmask;
        a.ops += sg.begin;
        a.n_ops = sg.end - sg.begin;
        a.optabs += sg.begin;
        a.optab_stride = c->P->n_ops;
    };
    for (int b0 = 0; b0 < batch; b0 += slice) {
        const int nb = std::min(slice, batch - b0);
        const double* dp = d_params + (size_t)b0 * c->n_params;
        if ((rc = run_tables(c, dp, nb, with_grad, st))) return rc;
        {
            dim3 grid(std::min(c->sm_count * 8, std::max(1, rows / 256)), nb);
            replicate_matrix<<<grid, 256, 0, st>>>(c->U.as<cplx>(), psi, rows);
            c->launches++;
        }
        if ((rc = run_optabs(c, nb, pf.log_ct, st))) return rc;
        if ((rc = run_dense_tabs(c, pf.log_ct, st))) return rc;
        time_begin(c, "fused_exec<WINDOW_FWD>", st);
        for (const auto& sg : c->segs) {
            ExecArgs a;
            seg_args(pf, sg, a);
            cudaError_t e = launch_fused_mode<MODE_APPLY>(a, pf, nb, st);
            c->launches++;
            if (e != cudaSuccess) return fail(SQGPU_ERR_CUDA, "fused_exec launch failed: %s", cudaGetErrorString(e));
        }
        time_end(c, st);
        {   // beta_N = conj(H psi_N); energy = Re <psi|H psi>
            launch_csr_matvec(rows, c->h_nnz, nb, c->hIndptr.as<int32_t>(), c->hIndices.as<int32_t>(), c->hValues.as<cplx>(), psi, lam, 1, st);
            dim3 g2(nblk, nb);
            expectation_partial<<<g2, 256, 0, st>>>(rows, psi, lam, -1.0, c->wTrPart.as<double>());
            sum_partials<<<nb, 32, 0, st>>>(c->wTrPart.as<double>(), nblk, 1, 1.0, d_energy + b0, 1);
            c->launches += 3;
            CUDA_TRY(cudaGetLastError());
        }
        if (!with_grad) continue;
        if (pb.log_ct != pf.log_ct && (rc = run_optabs(c, nb, pb.log_ct, st))) return rc;
        if ((rc = run_dense_tabs(c, pb.log_ct, st))) return rc;
        if (c->P->w_total > 0) CUDA_TRY(cudaMemsetAsync(c->wWPart.p, 0, (size_t)nb * pb.w_slices * c->P->w_total * sizeof(cplx), st));
        CUDA_TRY(cudaMemsetAsync(c->wTrPart.p, 0, (size_t)nb * pb.chunks * 6 * sizeof(double), st));
        time_begin(c, "fused_exec<WINDOW_BWD>", st);
        for (int si = (int)c->segs.size() - 1; si >= 0; --si) {
            ExecArgs a;
            seg_args(pb, c->segs[si], a);
            a.beta = lam;
            a.w_part = c->wWPart.as<cplx>();
            cudaError_t e = launch_fused_mode<MODE_BWD>(a, pb, nb, st);
            c->launches++;
            if (e != cudaSuccess) return fail(SQGPU_ERR_CUDA, "fused_exec launch failed: %s", cudaGetErrorString(e));
        }
        time_end(c, st);
        if (pb.w_slices > 1 && c->P->w_total > 0) {
            fold_w_chunks<<<dim3(fold_grid_x(c->P->w_total), nb), 256, 0, st>>>(c->wWPart.as<cplx>(), pb.w_slices, c->P->w_total);
            c->launches++;
        }
        reduce_partials<<<dim3(nb, reduce_grid_y(c->n_params)), 128, 0, st>>>(c->wTrPart.as<double>(), pb.chunks, c->wWPart.as<cplx>(), c->P->w_total, c->P->dOps.as<DevOp>(),
                                            c->P->dParamOp.as<int>(), c->P->dParamOp.as<int>() + std::max(c->n_params, 1), c->P->wDKtab.as<cplx>(),
                                            c->P->dkern_total, c->P->wKtab.as<cplx>(), c->P->kern_total, c->n_params, 1, c->wTraces.as<double>(), 1, pb.w_slices);
        grad_from_traces<<<nb, 128, 0, st>>>(c->wTraces.as<double>(), c->n_params, 2.0, d_grad + (size_t)b0 * c->n_params);
        c->launches += 2;
        CUDA_TRY(cudaGetLastError());
    }
    return SQGPU_OK;
}

static int vqe_dev(sqgpu_ctx* c, const double* d_params, int batch, bool with_grad, double* d_energy, double* d_grad, cudaStream_t st) {
    c->P = &c->plan2;
    int rc = check_ready(c, true);
    if (rc) return rc;
    if (batch < 0) return fail(SQGPU_ERR_INVALID, "negative batch");
    if (batch == 0) return SQGPU_OK;
    if (c->cols != 1) return fail(SQGPU_ERR_INVALID, "the VQE path needs a state vector (cols = 1), the resident matrix has %d columns", c->cols);
    if (!c->hIndptr.p || c->h_rows != c->rows) return fail(SQGPU_ERR_STATE, "no Hamiltonian of matching size set (call sqgpu_set_hamiltonian_csr)");
    if (!d_energy || (with_grad && !d_grad && c->n_params > 0)) return fail(SQGPU_ERR_INVALID, "NULL buffer");
    if (with_grad && !c->all_unitary) return fail(SQGPU_ERR_UNSUPPORTED, "gradient with a non-unitary GENERAL gate is not supported");
    {   // windowed shared-memory executor first; option vqe_stream (test hook) or a plan that does not fit: one op per launch
        if (!c->opt.vqe_stream) {
            rc = vqe_window_dev(c, d_params, batch, with_grad, d_energy, d_grad, st);
            if (rc != 1) return rc;
            c->P = &c->plan2;
        }
    }
    for (int k = 0; k < c->P->n_ops; ++k) {
        const DevOp& op = c->P->ops[k];
        const bool ok = op.dim == 2 || (op.dim == 4 && op.nq == 2 && op.ctrl_mask == 0);
        if (with_grad && !ok) return fail(SQGPU_ERR_UNSUPPORTED, "VQE gradient with 3+ qubit dense or multi-controlled gates is not implemented on the streaming path");
    }
    const int rows = c->rows;
    // parameter sets per slice: two state buffers of <= 1 GiB each
    const int slice = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::min(batch, 65535), ((size_t)1 << 30) / ((size_t)rows * sizeof(cplx))));
    const int nblk = std::min(c->sm_count * 8, std::max(1, rows / 512));
    if ((rc = c->wMat.ensure((size_t)2 * slice * rows * sizeof(cplx)))) return rc;
    if ((rc = c->wTrPart.ensure((size_t)slice * (nblk * 32 + 6) * sizeof(double)))) return rc;
    if (with_grad) {
        if ((rc = c->wWPart.ensure(std::max<size_t>(1, (size_t)slice * c->P->w_total) * sizeof(cplx)))) return rc;
        if ((rc = c->wTraces.ensure((size_t)slice * (1 + c->n_params) * 6 * sizeof(double)))) return rc;
    }
    cplx* psi = c->wMat.as<cplx>();
    cplx* lam = psi + (size_t)slice * rows;
    for (int b0 = 0; b0 < batch; b0 += slice) {
        const int nb = std::min(slice, batch - b0);
        const double* dp = d_params + (size_t)b0 * c->n_params;
        if ((rc = run_tables(c, dp, nb, with_grad, st))) return rc;
        {
            dim3 grid(std::min(c->sm_count * 8, std::max(1, rows / 256)), nb);
            replicate_matrix<<<grid, 256, 0, st>>>(c->U.as<cplx>(), psi, rows);
            c->launches++;
        }
        time_begin(c, "gate1q_stream", st);
        for (int k = 0; k < c->P->n_ops; ++k) {
            const DevOp& op = c->P->ops[k];
            const cplx* K = op.kern_off >= 0 ? c->P->wKtab.as<cplx>() + op.kern_off : c->dPool.as<cplx>() + op.pool_off;
            const long long kst = op.kern_off >= 0 ? c->P->kern_total : 0;
            if ((rc = launch_stream_gate(c, op, false, psi, rows, nb, rows, 1, 1, K, kst, st))) return rc;
        }
        time_end(c, st);
        {   // beta_N = conj(H psi_N); energy = Re <psi|H psi>
            launch_csr_matvec(rows, c->h_nnz, nb, c->hIndptr.as<int32_t>(), c->hIndices.as<int32_t>(), c->hValues.as<cplx>(), psi, lam, 1, st);
            dim3 g2(nblk, nb);
            expectation_partial<<<g2, 256, 0, st>>>(rows, psi, lam, -1.0, c->wTrPart.as<double>());
            sum_partials<<<nb, 32, 0, st>>>(c->wTrPart.as<double>(), nblk, 1, 1.0, d_energy + b0, 1);
            c->launches += 3;
            CUDA_TRY(cudaGetLastError());
        }
        if (!with_grad) continue;
        for (int k = c->P->n_ops - 1; k >= 0; --k) {
            const DevOp& op = c->P->ops[k];
            const cplx* K = op.kern_off >= 0 ? c->P->wKtab.as<cplx>() + op.kern_off : c->dPool.as<cplx>() + op.pool_off;
            const long long kst = op.kern_off >= 0 ? c->P->kern_total : 0;
            StreamGate g = make_stream_gate(op, psi, rows, rows, 1, 1, K, kst);
            const int want_w = op.n_params > 0 ? 1 : 0;
            int blocks, width;
            if (op.dim == 2) {
                const long long items = (long long)(rows >> g.nfix);
                blocks = (int)std::max<long long>(1, std::min<long long>((items + 255) / 256, nblk));
                width = 8;
                adjoint1q_stream<<<dim3(blocks, nb), 256, 0, st>>>(g, lam, rows, c->wTrPart.as<double>(), want_w);
            } else {
                const long long items = (long long)(rows >> 2);
                blocks = (int)std::max<long long>(1, std::min<long long>((items + 127) / 128, nblk));
                width = 32;
                adjoint2q_stream<<<dim3(blocks, nb), 128, 0, st>>>(g, lam, rows, c->wTrPart.as<double>(), want_w);
            }
            c->launches++;
            if (want_w) {
                sum_partials<<<nb, 32, 0, st>>>(c->wTrPart.as<double>(), blocks, width, 1.0,
                                                reinterpret_cast<double*>(c->wWPart.as<cplx>() + op.w_off), 2 * c->P->w_total);
                c->launches++;
            }
        }
        // dL_p = sum dK_p W  (same contraction as the unitary path), grad_p = 2 Re dL_p  (…Base.cpp:1180-1186)
        double* dummy_tr = c->wTrPart.as<double>() + (size_t)slice * nblk * 32;  // slice * 6 doubles, content unused
        reduce_partials<<<dim3(nb, reduce_grid_y(c->n_params)), 128, 0, st>>>(dummy_tr, 1, c->wWPart.as<cplx>(), c->P->w_total, c->P->dOps.as<DevOp>(), c->P->dParamOp.as<int>(),
                                            c->P->dParamOp.as<int>() + std::max(c->n_params, 1), c->P->wDKtab.as<cplx>(), c->P->dkern_total,
                                            c->P->wKtab.as<cplx>(), c->P->kern_total, c->n_params, 1, c->wTraces.as<double>());
        grad_from_traces<<<nb, 128, 0, st>>>(c->wTraces.as<double>(), c->n_params, 2.0, d_grad + (size_t)b0 * c->n_params);
        c->launches += 2;
        CUDA_TRY(cudaGetLastError());
    }
    return SQGPU_OK;
}
