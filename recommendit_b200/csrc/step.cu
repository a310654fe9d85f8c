// One-call fused training step (src/training/train_embeddings.py:183-192 of the reference).
// Pure orchestration: carves the workspace and enqueues the kernels of tower.cu / loss.cu /
// scatter_adam.cu on one stream.  No synchronisation, no allocation → CUDA-graph capturable.
#include "common.cuh"

int rb_scatter_tables(int phase, int n_tables, const int64_t* const ids_a[2], const int64_t* const ids_b[2], const int n_a[2],
                      const int n_b[2], const float* const rows[2], int D, const long long n_rows[2], long long padding_idx,
                      float* const dense[2], int64_t* const uniq_ids[2], float* const uniq_grads[2], int* const n_uniq[2],
                      int* const row_slot[2], void* workspace, size_t workspace_bytes, cudaStream_t st);
int rb_sumsq_accumulate(rb200_opt_state* st, const rb200_sumsq_seg* segs, int n_segs, int do_clip, void* workspace,
                        size_t workspace_bytes, cudaStream_t s);
int rb_tower_prep_tc(int n_sets, const float* const W1[], const float* const W2[], const int E[], int D, int H,
                     unsigned char* const img[], rb200_opt_state* opt, cudaStream_t st);
size_t rb_tower_img_bytes(int D, int H, int E);
bool rb_tower_tc_supported(int D, int H, int E);
int rb_adam_dense2(float* w0, const float* g0, float* m0, float* v0, long long n0, float* w1, const float* g1, float* m1, float* v1,
                   long long n1, const rb200_opt_state* st, cudaStream_t s);
int rb_adam_tables_dense2(float* w0, float* m0, float* v0, long long rows0, int* slot0, const float* ug0, float* w1, float* m1, float* v1,
                          long long rows1, int* slot1, const float* ug1, int D, const rb200_opt_state* st, cudaStream_t s);
int rb_adam_step_all(float* const mlp_w[2], const float* const mlp_g[2], float* const mlp_m[2], float* const mlp_v[2],
                     const long long mlp_n[2], float* const tab_w[2], float* const tab_m[2], float* const tab_v[2],
                     const long long tab_rows[2], int* const slot[2], const float* const ug[2], int D, rb200_opt_state* st,
                     const double* norm_partials, int n_norm_partials, cudaStream_t s);
int rb_norm_from_partials(const double* partials, int n, rb200_opt_state* st, cudaStream_t s);
int rb_bpr_pair(const float* u, const float* p, const float* n, int B, int D, float* loss, float* du, float* dp, float* dn,
                float grad_scale, float loss_scale, void* workspace, size_t workspace_bytes, rb200_opt_state* opt, cudaStream_t st,
                cudaStream_t st_fin, cudaEvent_t fork);
int rb_tower_bwd(const rb200_tower_bwd_job* jobs, int n_jobs, int D, int H, float dropout_p, int mode, float* grads_out,
                 int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t stream, RbPartials* defer);
int rb_grad_finish(int n_tables, const int64_t* const ids_a[2], const int64_t* const ids_b[2], const int n_a[2], const int n_b[2],
                   const float* const rows[2], int D, const long long n_rows[2], long long padding_idx, float* const dense[2],
                   int64_t* const uniq_ids[2], float* const uniq_grads[2], int* const n_uniq[2], int* const row_slot[2],
                   void* scatter_ws, size_t scatter_ws_bytes, const RbPartials red[2], float* const red_out[2], int do_sumsq,
                   void* sumsq_ws, size_t sumsq_ws_bytes, const double** norm_partials, int* n_norm_partials, cudaStream_t s);
int rb_sample_batch(const rb200_sampler& S, int B, long long epoch, long long step, const int64_t* counter_dev, int64_t* out_users,
                    int64_t* out_pos, int64_t* out_neg, cudaStream_t st);
int rb_reset_slots2(const int64_t* ids0, const int* n0, int cap0, int* slot0, const int64_t* ids1, const int* n1, int cap1, int* slot1,
                    cudaStream_t s);

namespace {

struct StepWs {
    float *u, *p, *n;              // tower outputs [B,D] each (p,n contiguous)
    float *hid_u, *hid_pn;         // [B,H], [2B,H]
    float *den_u, *den_pn;         // [B], [2B]
    float *du, *dpn;               // loss gradients [B,D], [2B,D]
    float *dpre_u, *dpre_pn, *dact_u, *dact_pn, *drows_u, *drows_pn;
    int64_t* ids_pn;               // [2B]
    float *g_user_mlp, *g_item_mlp;   // contiguous
    int64_t *uniq_u, *uniq_i; float *ug_u, *ug_i; int *n_uniq;   // n_uniq[0]=user, [1]=item
    unsigned char *img_user, *img_item;   // tensor-core weight images
    void *ws_bwd, *ws_bwd_u, *ws_loss, *ws_scatter, *ws_sumsq;
    size_t b_bwd, b_loss, b_scatter, b_sumsq;
    int P_user, P_item;
};

bool carve(RbArena& ar, const rb200_step_params& s, StepWs& w) {
    const size_t B = s.B, D = s.D, H = s.H;
    const int items = s.loss_kind == 0 ? 2 : 1;
    w.u = ar.take<float>(B * D);
    w.p = ar.take<float>(items * B * D); w.n = w.p + B * D;
    w.hid_u = ar.take<float>(B * H); w.hid_pn = ar.take<float>(items * B * H);
    w.den_u = ar.take<float>(B); w.den_pn = ar.take<float>(items * B);
    w.du = ar.take<float>(B * D); w.dpn = ar.take<float>(items * B * D);
    w.dpre_u = ar.take<float>(B * D); w.dpre_pn = ar.take<float>(items * B * D);
    w.dact_u = ar.take<float>(B * H); w.dact_pn = ar.take<float>(items * B * H);
    w.drows_u = ar.take<float>(B * D); w.drows_pn = ar.take<float>(items * B * D);
    w.ids_pn = ar.take<int64_t>(items * B);
    w.P_user = s.H * s.D + s.H + s.D * s.H + s.D;
    w.P_item = s.H * (s.D + s.extra_dim) + s.H + s.D * s.H + s.D;
    w.g_user_mlp = ar.take<float>((size_t)w.P_user + w.P_item);
    w.g_item_mlp = w.g_user_mlp + w.P_user;
    w.uniq_u = ar.take<int64_t>(B); w.uniq_i = ar.take<int64_t>(items * B);
    w.ug_u = ar.take<float>(B * D); w.ug_i = ar.take<float>(items * B * D);
    w.n_uniq = ar.take<int>(2);
    const bool tc = rb_tower_tc_supported(s.D, s.H, s.extra_dim);
    w.img_user = ar.take<unsigned char>(tc ? rb_tower_img_bytes(s.D, s.H, 0) : 16);
    w.img_item = ar.take<unsigned char>(tc ? rb_tower_img_bytes(s.D, s.H, s.extra_dim) : 16);
    w.b_bwd = rb200_tower_bwd_workspace_bytes(s.D, s.H, s.extra_dim);
    w.ws_bwd = ar.take<char>(w.b_bwd);
    w.ws_bwd_u = ar.take<char>(w.b_bwd);      // the user tower's backward runs concurrently with the item towers'
    const size_t l0 = rb200_bpr_pair_workspace_bytes(s.B), l1 = s.loss_kind == 1 ? rb200_bpr_inbatch_workspace_bytes(s.B, s.D) : 0;
    w.b_loss = l0 > l1 ? l0 : l1;
    w.ws_loss = ar.take<char>(w.b_loss);
    const size_t s0 = rb200_scatter_workspace_bytes(s.B, s.n_user_rows), s1 = rb200_scatter_workspace_bytes(items * s.B, s.n_item_rows);
    w.b_scatter = s0 + s1;
    w.ws_scatter = ar.take<char>(w.b_scatter);
    w.b_sumsq = rb200_sumsq_workspace_bytes();
    {   // fp64 block partials of the fused gradient-finish kernel (segment blocks + partial-reduction blocks)
        const size_t fin = 256 + sizeof(double) * ((size_t)(items + 1) * B / 8 + ((size_t)w.P_user + w.P_item) / 32 + 16);
        if (fin > w.b_sumsq) w.b_sumsq = fin;
    }
    w.ws_sumsq = ar.take<char>(w.b_sumsq);
    return ar.ok();
}

// Library-owned side stream + events for the fork/join inside the step (capturable: the side stream joins the capture
// through the event dependencies).
struct SideStream { cudaStream_t s = nullptr; cudaEvent_t fork = nullptr, fork2 = nullptr, fork3 = nullptr, join = nullptr, join2 = nullptr; };
int side_stream(SideStream** out) {
    static thread_local SideStream per_dev[64];
    int dev = 0;
    RB_CUDA(cudaGetDevice(&dev));
    RB_REQUIRE(dev >= 0 && dev < 64, "bad device index");
    SideStream& ss = per_dev[dev];
    if (!ss.s) {
        RB_CUDA(cudaStreamCreateWithFlags(&ss.s, cudaStreamNonBlocking));
        RB_CUDA(cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming));
        RB_CUDA(cudaEventCreateWithFlags(&ss.fork2, cudaEventDisableTiming));
        RB_CUDA(cudaEventCreateWithFlags(&ss.fork3, cudaEventDisableTiming));
        RB_CUDA(cudaEventCreateWithFlags(&ss.join2, cudaEventDisableTiming));
        RB_CUDA(cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming));
    }
    *out = &ss;
    return RB200_OK;
}

// (data-parallel mode: the loss is pre-scaled by 1/world like the gradients, so that one SUM all-reduce yields the mean)
__global__ void copy_loss_kernel(float* loss, rb200_opt_state* st, float scale) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { loss[0] *= scale; st->loss = loss[0]; }
}

int check(const rb200_step_params* s) {
    RB_REQUIRE(s, "bpr_step: NULL params");
    RB_REQUIRE(s->B >= 1 && s->extra_dim >= 0, "bpr_step: bad B/extra_dim");
    RB_REQUIRE(s->user_table && s->user_table_m && s->user_table_v && s->item_table && s->item_table_m && s->item_table_v,
               "bpr_step: NULL table pointer");
    RB_REQUIRE(s->user_mlp && s->user_mlp_m && s->user_mlp_v && s->item_mlp && s->item_mlp_m && s->item_mlp_v,
               "bpr_step: NULL MLP pointer");
    RB_REQUIRE(s->opt && s->loss && s->user_ids && s->pos_ids, "bpr_step: NULL opt/loss/ids");
    RB_REQUIRE(s->loss_kind == 1 || s->neg_ids, "bpr_step: neg_ids required for the pairwise loss");
    RB_REQUIRE(s->extra_dim == 0 || (s->pos_extra && (s->loss_kind == 1 || s->neg_extra)), "bpr_step: NULL extra");
    RB_REQUIRE(s->adam_mode == 1 || (s->user_row_slot && s->item_row_slot), "bpr_step: dense Adam needs row_slot buffers");
    RB_REQUIRE(s->workspace, "bpr_step: NULL workspace");
    return RB200_OK;
}

}  // namespace

extern "C" size_t rb200_bpr_step_workspace_bytes(int B, int D, int H, int extra_dim, int64_t n_user_rows, int64_t n_item_rows,
                                                 int loss_kind) {
    rb200_step_params s{};
    s.B = B; s.D = D; s.H = H; s.extra_dim = extra_dim; s.n_user_rows = n_user_rows; s.n_item_rows = n_item_rows;
    s.loss_kind = loss_kind;
    RbArena ar(nullptr, ~(size_t)0);
    StepWs w;
    carve(ar, s, w);
    return ar.off + 256;
}

extern "C" int rb200_bpr_step_views(const rb200_step_params* s, rb200_step_views* out) {
    int rc = check(s);
    if (rc) return rc;
    RB_REQUIRE(out, "bpr_step_views: NULL out");
    RbArena ar(s->workspace, s->workspace_bytes);
    StepWs w;
    if (!carve(ar, *s, w)) return rb_set_error(RB200_ERR_WORKSPACE, "bpr_step_views: workspace too small");
    out->user_mlp_grad = w.g_user_mlp; out->item_mlp_grad = w.g_item_mlp;
    out->user_uniq_ids = w.uniq_u; out->item_uniq_ids = w.uniq_i;
    out->user_uniq_grads = w.ug_u; out->item_uniq_grads = w.ug_i;
    out->user_n_uniq = w.n_uniq; out->item_n_uniq = w.n_uniq + 1;
    out->user_emb = w.u; out->pos_emb = w.p; out->neg_emb = s->loss_kind == 0 ? w.n : nullptr;
    return RB200_OK;
}

extern "C" int rb200_bpr_step(const rb200_step_params* s, void* stream) {
    int rc = check(s);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    RbArena ar(s->workspace, s->workspace_bytes);
    StepWs w;
    if (!carve(ar, *s, w)) return rb_set_error(RB200_ERR_WORKSPACE, "bpr_step: workspace too small (%zu given, %zu needed)",
                                               s->workspace_bytes, ar.off);
    const int B = s->B, D = s->D, H = s->H, E = s->extra_dim;
    const bool pair = s->loss_kind == 0;
    const int items = pair ? 2 : 1;
    const int Pu = w.P_user, Pi = w.P_item;
    const int Din_i = D + E;
    if (s->dp_grads) { w.g_user_mlp = s->dp_grads; w.g_item_mlp = s->dp_grads + Pu; }
    const float gscale = s->grad_scale > 0.f ? s->grad_scale : 1.f;

    int ev_i = 0;
#define RB_STAGE_EVENT()                                                                   \
    do {                                                                                   \
        if (s->stage_events_host) RB_CUDA(cudaEventRecord((cudaEvent_t)s->stage_events_host[ev_i], st)); \
        ++ev_i;                                                                            \
    } while (0)
    RB_STAGE_EVENT();
    const bool tc = s->tower_mode != 0;
    if (!tc && (rc = rb200_opt_begin_step(s->opt, st))) return rc;      // (tensor-core modes: done by the weight-image kernel)

    // ---- fork: the (id, sample) sort needs only the ids, so it runs on a side stream under the towers ------- //
    const bool dense = s->adam_mode == 0;
    const int64_t* sc_ia[2] = {s->user_ids, s->pos_ids};
    const int64_t* sc_ib[2] = {nullptr, pair ? s->neg_ids : nullptr};
    const int sc_na[2] = {B, B}, sc_nb[2] = {0, pair ? B : 0};
    const float* sc_rw[2] = {w.drows_u, w.drows_pn};
    const long long sc_nr[2] = {s->n_user_rows, s->n_item_rows};
    const bool dp = s->dp_grads != nullptr;
    float* dp_user_tab = dp ? s->dp_grads + (size_t)w.P_user + w.P_item : nullptr;
    float* dp_item_tab = dp ? dp_user_tab + (size_t)s->n_user_rows * D : nullptr;
    float* sc_dn[2] = {dp_user_tab, dp_item_tab};
    int64_t* sc_ui[2] = {w.uniq_u, w.uniq_i};
    float* sc_ug[2] = {w.ug_u, w.ug_i};
    int* sc_nu[2] = {w.n_uniq, w.n_uniq + 1};
    int* sc_rs[2] = {(dense && !dp) ? s->user_row_slot : nullptr, (dense && !dp) ? s->item_row_slot : nullptr};
    if (dp) RB_CUDA(cudaMemsetAsync(dp_user_tab, 0, sizeof(float) * (size_t)(s->n_user_rows + s->n_item_rows) * D, st));
    const bool fast_scatter = items * B <= 16384 && s->n_user_rows < (1ll << 31) && s->n_item_rows < (1ll << 31);
    SideStream* side = nullptr;
    if ((rc = side_stream(&side))) return rc;
    if (fast_scatter) {
        RB_CUDA(cudaEventRecord(side->fork, st));
        RB_CUDA(cudaStreamWaitEvent(side->s, side->fork, 0));
        if ((rc = rb_scatter_tables(1, 2, sc_ia, sc_ib, sc_na, sc_nb, sc_rw, D, sc_nr, s->padding_idx, sc_dn, sc_ui, sc_ug, sc_nu,
                                    sc_rs, w.ws_scatter, w.b_scatter, side->s))) return rc;
    }

    // ---- tensor-core modes: stage both towers' weight images once for forward and backward ----------- //
    if (tc) {
        RB_REQUIRE(rb_tower_tc_supported(D, H, E), "bpr_step: tower_mode %d needs D in {64,128}, H=128, extra_dim<=24", s->tower_mode);
        const float* pw1[2] = {s->user_mlp, s->item_mlp};
        const float* pw2[2] = {s->user_mlp + H * D + H, s->item_mlp + H * Din_i + H};
        const int pe[2] = {0, E};
        unsigned char* pim[2] = {w.img_user, w.img_item};
        if ((rc = rb_tower_prep_tc(2, pw1, pw2, pe, D, H, pim, s->opt, st))) return rc;
    }

    // ---- forward: user / positive / negative towers in one launch -------------------------- //
    rb200_tower_job fj[3] = {};
    fj[0].img = tc ? w.img_user : nullptr;
    fj[0].table = s->user_table; fj[0].ids = s->user_ids; fj[0].extra = nullptr; fj[0].extra_dim = 0;
    fj[0].W1 = s->user_mlp; fj[0].b1 = s->user_mlp + H * D; fj[0].W2 = s->user_mlp + H * D + H; fj[0].b2 = s->user_mlp + H * D + H + D * H;
    fj[0].out = w.u; fj[0].hid = w.hid_u; fj[0].denom = w.den_u; fj[0].keep_mask = s->keep_mask_user;
    fj[0].n_rows = s->n_user_rows; fj[0].B = B;
    for (int t = 0; t < items; ++t) {
        rb200_tower_job& j = fj[1 + t];
        j.table = s->item_table; j.ids = t == 0 ? s->pos_ids : s->neg_ids; j.extra = t == 0 ? s->pos_extra : s->neg_extra;
        j.extra_dim = E; j.extra_by_id = s->extra_by_id;
        j.W1 = s->item_mlp; j.b1 = s->item_mlp + H * Din_i; j.W2 = s->item_mlp + H * Din_i + H; j.b2 = s->item_mlp + H * Din_i + H + D * H;
        j.out = w.p + (size_t)t * B * D; j.hid = w.hid_pn + (size_t)t * B * H; j.denom = w.den_pn + (size_t)t * B;
        j.keep_mask = t == 0 ? s->keep_mask_pos : s->keep_mask_neg;
        j.n_rows = s->n_item_rows; j.B = B;
        j.img = tc ? w.img_item : nullptr;
    }
    if ((rc = rb200_tower_fwd(fj, 1 + items, D, H, s->dropout_p, s->seed, 0, &s->opt->step, s->tower_mode, s->err_flag, nullptr, 0,
                              st))) return rc;

    RB_STAGE_EVENT();
    // ---- loss + gradient w.r.t. the tower outputs ------------------------------------------- //
    // (pairwise: the scalar reduction of the loss goes to the side stream — nothing downstream needs it before the join)
    if (pair) {
        rc = rb_bpr_pair(w.u, w.p, w.n, B, D, s->loss, w.du, w.dpn, w.dpn + (size_t)B * D, gscale, dp ? gscale : 1.f, w.ws_loss, w.b_loss,
                         s->opt, st, side->s, side->fork2);
        if (rc) return rc;
    } else {
        if ((rc = rb200_bpr_inbatch(w.u, w.p, B, D, s->inbatch_mode, s->loss, w.du, w.dpn, gscale, w.ws_loss, w.b_loss, st))) return rc;
        copy_loss_kernel<<<1, 32, 0, st>>>(s->loss, s->opt, dp ? gscale : 1.f);
        RB_LAUNCH_CHECK("copy_loss_kernel");
        RB_CUDA(cudaEventRecord(side->fork2, st));
        RB_CUDA(cudaStreamWaitEvent(side->s, side->fork2, 0));
    }

    RB_STAGE_EVENT();
    // ---- backward through the towers ----------------------------------------------------------- //
    rb200_tower_bwd_job bj[2] = {};
    bj[0].table = s->user_table; bj[0].ids = s->user_ids; bj[0].extra = nullptr; bj[0].n_rows = s->n_user_rows; bj[0].B = B;
    bj[0].extra_dim = 0; bj[0].W1 = fj[0].W1; bj[0].W2 = fj[0].W2; bj[0].dY = w.du; bj[0].y = w.u; bj[0].denom = w.den_u;
    bj[0].hid = w.hid_u; bj[0].dpre = w.dpre_u; bj[0].dact = w.dact_u; bj[0].dRows = w.drows_u;
    bj[0].img = tc ? w.img_user : nullptr;
    // (second fork, recorded with the loss above): the user tower's backward goes to the side stream (behind the id sort, long finished by now) and
    // overlaps the item towers' backward — neither fills the GPU on its own (64 / 128 tiles of 128 samples at B = 8192)
    // tensor-core modes + in-graph scatter: the split-K partials stay unreduced; rb_grad_finish reduces them together with the
    // segment sums and the gradient norm in one launch
    const bool fuse_finish = tc && fast_scatter;
    RbPartials red[2] = {};
    const double* norm_part = nullptr;           // Σg² block partials of the fused finish, finalised inside the Adam launch
    int n_norm_part = 0;
    if ((rc = rb_tower_bwd(bj, 1, D, H, s->dropout_p, s->tower_mode, w.g_user_mlp, 0, w.ws_bwd_u, w.b_bwd, side->s,
                           fuse_finish ? &red[0] : nullptr))) return rc;
    RB_CUDA(cudaEventRecord(side->join, side->s));
    for (int t = 0; t < items; ++t) {
        rb200_tower_bwd_job& j = bj[t];
        j = rb200_tower_bwd_job{};
        j.table = s->item_table; j.ids = fj[1 + t].ids; j.extra = fj[1 + t].extra; j.n_rows = s->n_item_rows; j.B = B;
        j.extra_dim = E; j.extra_by_id = s->extra_by_id; j.W1 = fj[1].W1; j.W2 = fj[1].W2;
        j.dY = w.dpn + (size_t)t * B * D; j.y = w.p + (size_t)t * B * D; j.denom = w.den_pn + (size_t)t * B;
        j.hid = w.hid_pn + (size_t)t * B * H; j.dpre = w.dpre_pn + (size_t)t * B * D; j.dact = w.dact_pn + (size_t)t * B * H;
        j.dRows = w.drows_pn + (size_t)t * B * D;
        j.img = tc ? w.img_item : nullptr;
    }
    if ((rc = rb_tower_bwd(bj, items, D, H, s->dropout_p, s->tower_mode, w.g_item_mlp, 0, w.ws_bwd, w.b_bwd, st,
                           fuse_finish ? &red[1] : nullptr))) return rc;

    RB_STAGE_EVENT();
    // ---- sparse embedding gradients: deterministic sorted-segment sums ---------------------- //
    RB_CUDA(cudaStreamWaitEvent(st, side->join, 0));          // join: user-tower gradients and sorted positions / segment starts
    if (fuse_finish) {
        float* red_out[2] = {w.g_user_mlp, w.g_item_mlp};
        if ((rc = rb_grad_finish(2, sc_ia, sc_ib, sc_na, sc_nb, sc_rw, D, sc_nr, s->padding_idx, sc_dn, sc_ui, sc_ug, sc_nu, sc_rs,
                                 w.ws_scatter, w.b_scatter, red, red_out, dp ? 0 : 1, w.ws_sumsq, w.b_sumsq, &norm_part, &n_norm_part,
                                 st))) return rc;
    } else if (fast_scatter) {
        if ((rc = rb_scatter_tables(2, 2, sc_ia, sc_ib, sc_na, sc_nb, sc_rw, D, sc_nr, s->padding_idx, sc_dn, sc_ui, sc_ug, sc_nu,
                                    sc_rs, w.ws_scatter, w.b_scatter, st))) return rc;
    } else {
        RB_CUDA(cudaMemcpyAsync(w.ids_pn, s->pos_ids, sizeof(int64_t) * B, cudaMemcpyDeviceToDevice, st));
        if (pair) RB_CUDA(cudaMemcpyAsync(w.ids_pn + B, s->neg_ids, sizeof(int64_t) * B, cudaMemcpyDeviceToDevice, st));
        if ((rc = rb200_scatter_rows(s->user_ids, w.drows_u, B, D, s->n_user_rows, s->padding_idx, nullptr, w.uniq_u, w.ug_u,
                                     w.n_uniq, sc_rs[0], w.ws_scatter, w.b_scatter, st))) return rc;
        if ((rc = rb200_scatter_rows(w.ids_pn, w.drows_pn, items * B, D, s->n_item_rows, s->padding_idx, nullptr, w.uniq_i,
                                     w.ug_i, w.n_uniq + 1, sc_rs[1], w.ws_scatter, w.b_scatter, st))) return rc;
    }

    // ---- device-side batch producer: the ids of this step are dead from here on (sort, gathers and the weight-gradient
    //      kernels have run), so the NEXT batch is sampled into the same buffers on the side stream, under the optimizer -- //
    bool next_pending = false;
    if (s->next_batch) {
        RB_REQUIRE(pair, "bpr_step: next_batch produces (user, positive, negative) triples: loss_kind must be 0");
        RB_REQUIRE(E == 0 || s->extra_by_id, "bpr_step: next_batch needs the genre table mode (extra_by_id)");
        RB_CUDA(cudaEventRecord(side->fork3, st));
        RB_CUDA(cudaStreamWaitEvent(side->s, side->fork3, 0));
        if ((rc = rb_sample_batch(*s->next_batch, B, 0, 0, reinterpret_cast<const int64_t*>(&s->opt->step),
                                  const_cast<int64_t*>(s->user_ids), const_cast<int64_t*>(s->pos_ids), const_cast<int64_t*>(s->neg_ids),
                                  side->s))) return rc;
        RB_CUDA(cudaEventRecord(side->join2, side->s));
        next_pending = true;
    }
    if (dp) {                     // gradients are complete and dense in dp_grads: the caller all-reduces, then rb200_bpr_apply
        if (next_pending) RB_CUDA(cudaStreamWaitEvent(st, side->join2, 0));
        return RB200_OK;
    }

    RB_STAGE_EVENT();
    // ---- clip_grad_norm_(all parameters, 1.0) ---------------------------------------------------- //
    rb200_sumsq_seg segs[3] = {
        {w.g_user_mlp, (int64_t)Pu + Pi, nullptr, 0},
        {w.ug_u, (int64_t)B * D, w.n_uniq, D},
        {w.ug_i, (int64_t)items * B * D, w.n_uniq + 1, D},
    };
    if (!fuse_finish && (rc = rb_sumsq_accumulate(s->opt, segs, 3, 1, w.ws_sumsq, w.b_sumsq, st))) return rc;

    RB_STAGE_EVENT();
    // ---- Adam ---------------------------------------------------------------------------------- //
    int fused = 1;
    if (dense) {
        float* mw[2] = {s->user_mlp, s->item_mlp}; const float* mg[2] = {w.g_user_mlp, w.g_item_mlp};
        float* mm[2] = {s->user_mlp_m, s->item_mlp_m}; float* mv[2] = {s->user_mlp_v, s->item_mlp_v};
        const long long mn[2] = {Pu, Pi};
        float* tw[2] = {s->user_table, s->item_table}; float* tm[2] = {s->user_table_m, s->item_table_m};
        float* tv[2] = {s->user_table_v, s->item_table_v}; const long long tr[2] = {s->n_user_rows, s->n_item_rows};
        int* ts[2] = {s->user_row_slot, s->item_row_slot}; const float* tg[2] = {w.ug_u, w.ug_i};
        fused = rb_adam_step_all(mw, mg, mm, mv, mn, tw, tm, tv, tr, ts, tg, D, s->opt, norm_part, n_norm_part, st);
        if (fused < 0) return fused;
    }
    if (norm_part && fused != 0 && (rc = rb_norm_from_partials(norm_part, n_norm_part, s->opt, st))) return rc;
    if (fused == 1 && (rc = rb_adam_dense2(s->user_mlp, w.g_user_mlp, s->user_mlp_m, s->user_mlp_v, Pu, s->item_mlp, w.g_item_mlp,
                                           s->item_mlp_m, s->item_mlp_v, Pi, s->opt, st))) return rc;
    if (dense && fused == 1) {
        if ((rc = rb_adam_tables_dense2(s->user_table, s->user_table_m, s->user_table_v, s->n_user_rows, s->user_row_slot, w.ug_u,
                                        s->item_table, s->item_table_m, s->item_table_v, s->n_item_rows, s->item_row_slot, w.ug_i, D,
                                        s->opt, st))) return rc;
        if ((rc = rb_reset_slots2(w.uniq_u, w.n_uniq, B, s->user_row_slot, w.uniq_i, w.n_uniq + 1, items * B, s->item_row_slot, st)))
            return rc;
    } else if (!dense) {
        if ((rc = rb200_adam_rows(s->user_table, s->user_table_m, s->user_table_v, D, w.uniq_u, w.ug_u, w.n_uniq, B, s->opt, st))) return rc;
        if ((rc = rb200_adam_rows(s->item_table, s->item_table_m, s->item_table_v, D, w.uniq_i, w.ug_i, w.n_uniq + 1, items * B,
                                  s->opt, st))) return rc;
    }
    if (next_pending) RB_CUDA(cudaStreamWaitEvent(st, side->join2, 0));      // the side stream rejoins (graph capture needs it)
    RB_STAGE_EVENT();
#undef RB_STAGE_EVENT
    return RB200_OK;
}


extern "C" size_t rb200_bpr_dp_grad_floats(int D, int H, int extra_dim, int64_t n_user_rows, int64_t n_item_rows) {
    const size_t Pu = (size_t)H * D + H + (size_t)D * H + D, Pi = (size_t)H * (D + extra_dim) + H + (size_t)D * H + D;
    return Pu + Pi + (size_t)(n_user_rows + n_item_rows) * D;
}

extern "C" int rb200_bpr_apply(const rb200_step_params* s, void* stream) {
    int rc = check(s);
    if (rc) return rc;
    RB_REQUIRE(s->dp_grads, "bpr_apply: dp_grads is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    RbArena ar(s->workspace, s->workspace_bytes);
    StepWs w;
    if (!carve(ar, *s, w)) return rb_set_error(RB200_ERR_WORKSPACE, "bpr_apply: workspace too small");
    const int D = s->D;
    const size_t Pu = w.P_user, Pi = w.P_item;
    const size_t nu = (size_t)s->n_user_rows * D, ni = (size_t)s->n_item_rows * D;
    float* g = s->dp_grads;
    rb200_sumsq_seg seg[1] = {{g, (int64_t)(Pu + Pi + nu + ni), nullptr, 0}};
    if ((rc = rb_sumsq_accumulate(s->opt, seg, 1, 1, w.ws_sumsq, w.b_sumsq, st))) return rc;
    if ((rc = rb_adam_dense2(s->user_mlp, g, s->user_mlp_m, s->user_mlp_v, (long long)Pu, s->item_mlp, g + Pu, s->item_mlp_m,
                             s->item_mlp_v, (long long)Pi, s->opt, st))) return rc;
    if ((rc = rb_adam_dense2(s->user_table, g + Pu + Pi, s->user_table_m, s->user_table_v, (long long)nu, s->item_table,
                             g + Pu + Pi + nu, s->item_table_m, s->item_table_v, (long long)ni, s->opt, st))) return rc;
    return RB200_OK;
}
