// group.inl -- ONE process driving the ADMM loop on SEVERAL GPUs (included at the end of mg.cu, after admm.inl).
//
// The reference spreads the bodies of MCONTACT::CONTACT_ANALYSIS over host threads (`#pragma omp parallel for`,
// MCONTACT.h:2511,2629,2689).  A group is the same idea with devices: every member is an ordinary ddpca_admm handle
// on its own GPU holding a subset of the bodies (ddpca_admm_set_partition); the group issues the phases of the loop
// body on all members (asynchronously, one host thread) and moves the three small exchanges of an iteration itself,
// over NVLink peer copies ordered by events -- no second process, no NCCL:
//   * coarse right-hand side: every member's part is copied to member 0, summed there in member order, copied back;
//   * interface traces: the member that produced a signed side trace copies it straight into the receive buffer of
//     the member that owns the other side (pairwise, cudaMemcpyPeerAsync on the producer's stream);
//   * MONITOR sums: copied to pinned host memory and summed in member order (slots of remote bodies are zero).
// Results do not depend on the number of devices (per-body arithmetic is independent of the batch it runs in, the
// sums above add exact zeros or run in a fixed order) except for the coarse right-hand side, whose partial sums
// follow the partition.

struct ddpca_admm_group {
    std::vector<ddpca_admm *> m;
    std::vector<int> dev;
    int nb = 0, ni = 0, muscSett = 0;
    std::vector<int> body_rank;
    std::vector<double *> glob, send, recv, moni;        // per member, on its device
    std::vector<cudaEvent_t> ev_a, ev_b;                  // per member: "my part is ready" / "the sum is ready" etc.
    double *stage0 = nullptr;                             // member 0: [(n members + 1) x nglob] parts of the reduction + its result
    std::vector<double *> moni_host;                      // pinned, per member
    long nglob = 0, nmoni = 0;
    bool finalized = false;
    long cg_iters = 0;
    double cg_dof_iters = 0;
};

// out[i] = sum_k part[k * n + i], k ascending (fixed order)
__global__ void k_sum_members(int nmem, long n, const double *__restrict__ part, double *__restrict__ out)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int k = 0; k < nmem; k++) s += part[(size_t)k * n + i];
    out[i] = s;
}

static void group_free(ddpca_admm_group *g)
{
    if (!g) return;
    for (size_t k = 0; k < g->m.size(); k++) {
        cudaSetDevice(g->dev[k]);
        if (k < g->glob.size()) cudaFree(g->glob[k]);
        if (k < g->send.size()) cudaFree(g->send[k]);
        if (k < g->recv.size()) cudaFree(g->recv[k]);
        if (k < g->moni.size()) cudaFree(g->moni[k]);
        if (k < g->moni_host.size() && g->moni_host[k]) cudaFreeHost(g->moni_host[k]);
        if (k < g->ev_a.size() && g->ev_a[k]) cudaEventDestroy(g->ev_a[k]);
        if (k < g->ev_b.size() && g->ev_b[k]) cudaEventDestroy(g->ev_b[k]);
        if (k == 0) cudaFree(g->stage0);
        admm_free(g->m[k]);
    }
    delete g;
}

// sum of the members' coarse right-hand sides: to member 0, fixed-order sum, back to everybody
static int group_allreduce_glob(ddpca_admm_group *g, long n)
{
    const int nm = (int)g->m.size();
    if (nm == 1 || n == 0) return 0;
    for (int k = 0; k < nm; k++) {
        CU(cudaSetDevice(g->dev[k]));
        CU(cudaMemcpyPeerAsync(g->stage0 + (size_t)k * n, g->dev[0], g->glob[k], g->dev[k], sizeof(double) * n, g->m[k]->stream));
        CU(cudaEventRecord(g->ev_a[k], g->m[k]->stream));
    }
    // The sum goes to its own slot behind the members' parts: the members (0 included) copy it from there, so nobody
    // reads a buffer another member is already modifying (the interface-eliminated problem adds its constant part to the
    // right-hand side in place).  The slot is rewritten by the next sum only, which waits for every member's next part,
    // i.e. for their copies of this one (stream order).
    double *res = g->stage0 + (size_t)nm * g->nglob;
    CU(cudaSetDevice(g->dev[0]));
    for (int k = 1; k < nm; k++) CU(cudaStreamWaitEvent(g->m[0]->stream, g->ev_a[k], 0));
    k_sum_members<<<cdiv(n, 256), 256, 0, g->m[0]->stream>>>(nm, n, g->stage0, res);
    g->m[0]->launches++;
    CU(cudaEventRecord(g->ev_b[0], g->m[0]->stream));
    for (int k = 0; k < nm; k++) {
        CU(cudaSetDevice(g->dev[k]));
        if (k) CU(cudaStreamWaitEvent(g->m[k]->stream, g->ev_b[0], 0));
        CU(cudaMemcpyPeerAsync(g->glob[k], g->dev[k], res, g->dev[0], sizeof(double) * n, g->m[k]->stream));
    }
    return 0;
}

extern "C" {

int ddpca_partition_bodies(int nbody, const double *weight, int niface, const int *contBody, int nranks, int *body_rank)
{
    // greedy bin packing by weight, heaviest first; a body joins a rank that already holds a neighbour when that keeps
    // the rank within 5 % of a perfect share above the lightest one (NVSwitch: all pairs equidistant, only load matters)
    if (nbody < 1 || !weight || nranks < 1 || !body_rank || (niface > 0 && !contBody)) return fail("ddpca_partition_bodies: bad argument");
    std::vector<int> order(nbody);
    for (int v = 0; v < nbody; v++) { order[v] = v; body_rank[v] = -1; }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return weight[a] > weight[b]; });
    std::vector<double> load(nranks, 0.0);
    double total = 0.0;
    for (int v = 0; v < nbody; v++) total += weight[v];
    std::vector<std::vector<int>> neigh(nbody);
    for (int ts = 0; ts < niface; ts++) {
        const int a = contBody[2 * ts], b = contBody[2 * ts + 1];
        if (a < 0 || a >= nbody || b < 0 || b >= nbody) return fail("ddpca_partition_bodies: contBody out of range");
        neigh[a].push_back(b); neigh[b].push_back(a);
    }
    for (int v : order) {
        int best = 0;
        for (int r = 1; r < nranks; r++) if (load[r] < load[best]) best = r;
        int pick = best;
        double pick_load = 1e300;
        for (int u : neigh[v]) {
            const int r = body_rank[u];
            if (r < 0) continue;
            if (load[r] + weight[v] <= load[best] + weight[v] + 0.05 * total / nranks && load[r] < pick_load) { pick = r; pick_load = load[r]; }
        }
        body_rank[v] = pick;
        load[pick] += weight[v];
    }
    return 0;
}

int ddpca_admm_group_create(int ndev, const int *devices, int nbody, int niface, int muscSett, const int *body_rank, ddpca_admm_group **out)
{
    if (!out || ndev < 1 || !devices || !body_rank) return fail("ddpca_admm_group_create: bad argument");
    for (int v = 0; v < nbody; v++) if (body_rank[v] < 0 || body_rank[v] >= ndev) return fail("ddpca_admm_group_create: body_rank out of range");
    ddpca_admm_group *g = new ddpca_admm_group();
    g->nb = nbody; g->ni = niface; g->muscSett = muscSett;
    g->body_rank.assign(body_rank, body_rank + nbody);
    for (int k = 0; k < ndev; k++) {
        ddpca_admm *h = nullptr;
        if (ddpca_admm_create(devices[k], nbody, niface, muscSett, &h) || ddpca_admm_set_partition(h, body_rank, k)) { if (h) admm_free(h); group_free(g); return 1; }
        g->m.push_back(h);
        g->dev.push_back(devices[k]);
    }
    // direct NVLink copies between the members
    for (int a = 0; a < ndev; a++)
        for (int b = 0; b < ndev; b++) {
            if (a == b || devices[a] == devices[b]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[a], devices[b]);
            if (can) { cudaSetDevice(devices[a]); cudaDeviceEnablePeerAccess(devices[b], 0); cudaGetLastError(); }
        }
    *out = g;
    return 0;
}
int ddpca_admm_group_size(const ddpca_admm_group *g) { return g ? (int)g->m.size() : -1; }
ddpca_admm *ddpca_admm_group_member(ddpca_admm_group *g, int k) { return (g && k >= 0 && k < (int)g->m.size()) ? g->m[k] : nullptr; }
int ddpca_admm_group_owner(const ddpca_admm_group *g, int v) { return (g && v >= 0 && v < g->nb) ? g->body_rank[v] : -1; }
int ddpca_admm_group_device(const ddpca_admm_group *g, int k) { return (g && k >= 0 && k < (int)g->m.size()) ? g->dev[k] : -1; }
int ddpca_admm_group_destroy(ddpca_admm_group *g) { group_free(g); return 0; }

int ddpca_admm_group_finalize(ddpca_admm_group *g)
{
    if (!g) return fail("null group");
    if (g->finalized) return 0;
    const int nm = (int)g->m.size();
    g->glob.assign(nm, nullptr); g->send.assign(nm, nullptr); g->recv.assign(nm, nullptr); g->moni.assign(nm, nullptr);
    g->moni_host.assign(nm, nullptr); g->ev_a.assign(nm, nullptr); g->ev_b.assign(nm, nullptr);
    for (int k = 0; k < nm; k++) {
        CU(cudaSetDevice(g->dev[k]));
        long ng = 0, nt = 0, nmo = 0;
        if (ddpca_admm_exchange_sizes(g->m[k], &ng, &nt, &nmo)) return 1;
        g->nglob = ng; g->nmoni = nmo;
        if (dev_vec(nullptr, ng, &g->glob[k]) || dev_vec(nullptr, nt, &g->send[k]) || dev_vec(nullptr, nt, &g->recv[k]) || dev_vec(nullptr, nmo, &g->moni[k])) return 1;
        CU(cudaMallocHost(&g->moni_host[k], sizeof(double) * std::max<long>(1, nmo)));
        CU(cudaEventCreateWithFlags(&g->ev_a[k], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&g->ev_b[k], cudaEventDisableTiming));
        if (nm > 1 && ddpca_admm_set_exchange(g->m[k], g->glob[k], g->send[k], g->recv[k], g->moni[k])) return 1;
        if (ddpca_admm_finalize(g->m[k])) { g_err = "member " + std::to_string(k) + ": " + g_err; return 1; }
    }
    if (nm > 1) { CU(cudaSetDevice(g->dev[0])); if (dev_vec(nullptr, (long)(nm + 1) * g->nglob, &g->stage0)) return 1; }
    g->finalized = true;
    return 0;
}

int ddpca_admm_group_step(ddpca_admm_group *g, int apply_macro, double *monitor_row, long *cg_iters, double *cg_dof_iters)
{
    if (!g || !g->finalized) return fail("ddpca_admm_group_step: group not finalized");
    const int nm = (int)g->m.size();
    if (nm == 1) return ddpca_admm_step(g->m[0], apply_macro, monitor_row, cg_iters, cg_dof_iters);
#define EACH(call) for (int k = 0; k < nm; k++) { CU(cudaSetDevice(g->dev[k])); if (call(g->m[k])) return 1; }
    EACH(admm_bodies);
    if (apply_macro && (g->muscSett & 1)) {
        EACH(admm_macro_partial);
        if (group_allreduce_glob(g, g->m[0]->nglob)) return 1;
        EACH(admm_macro_apply);
    }
    if (apply_macro && (g->muscSett & 2)) {
        EACH(admm_macro1_partial);
        if (group_allreduce_glob(g, g->m[0]->nglob1)) return 1;
        EACH(admm_macro1_apply);
    }
    EACH(admm_traces);
    // pairwise swap: member k writes its range for peer p into p's receive range for peer k
    for (int k = 0; k < nm; k++) {
        ddpca_admm *a = g->m[k];
        CU(cudaSetDevice(g->dev[k]));
        for (size_t i = 0; i < a->peers.size(); i++) {
            const int p = a->peers[i];
            ddpca_admm *b = g->m[p];
            size_t j = 0;
            while (j < b->peers.size() && b->peers[j] != k) j++;
            const long cnt = a->peer_off[i + 1] - a->peer_off[i];
            if (j == b->peers.size() || b->peer_off[j + 1] - b->peer_off[j] != cnt) return fail("ddpca_admm_group_step: the members disagree on their shared interfaces");
            if (cnt) CU(cudaMemcpyPeerAsync(g->recv[p] + b->peer_off[j], g->dev[p], g->send[k] + a->peer_off[i], g->dev[k], sizeof(double) * cnt, a->stream));
        }
        CU(cudaEventRecord(g->ev_a[k], a->stream));
    }
    for (int k = 0; k < nm; k++) {
        CU(cudaSetDevice(g->dev[k]));
        for (int p : g->m[k]->peers) CU(cudaStreamWaitEvent(g->m[k]->stream, g->ev_a[p], 0));
    }
    EACH(admm_interface);
    EACH(admm_monitor);
#undef EACH
    // MONITOR sums: member order, on the host
    for (int k = 0; k < nm; k++) {
        CU(cudaSetDevice(g->dev[k]));
        CU(cudaMemcpyAsync(g->moni_host[k], g->moni[k], sizeof(double) * g->nmoni, cudaMemcpyDeviceToHost, g->m[k]->stream));
    }
    g->cg_iters = 0; g->cg_dof_iters = 0;
    for (int k = 0; k < nm; k++) {
        ddpca_admm *h = g->m[k];
        CU(cudaSetDevice(g->dev[k]));
        CU(cudaStreamSynchronize(h->stream));
        CU(cudaGetLastError());
        if (admm_bodies_finish(h)) return 1;
        if (!h->launch_err.empty()) { std::string msg = h->launch_err; h->launch_err.clear(); return fail(msg); }
        g->cg_iters += h->cg_iters; g->cg_dof_iters += h->cg_dof_iters;
    }
    ddpca_admm *h0 = g->m[0];
    for (long i = 0; i < g->nmoni; i++) {
        double s = 0.0;
        for (int k = 0; k < nm; k++) s += g->moni_host[k][i];
        h0->moni_host[i] = s;
    }
    admm_format_row(h0, monitor_row);
    if (cg_iters) *cg_iters = g->cg_iters;
    if (cg_dof_iters) *cg_dof_iters = g->cg_dof_iters;
    return 0;
}

long ddpca_admm_group_launch_count(ddpca_admm_group *g, int reset)
{
    if (!g) return -1;
    long n = 0;
    for (ddpca_admm *h : g->m) n += ddpca_admm_launch_count(h, reset);
    return n;
}

}  // extern "C"
