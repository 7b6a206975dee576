// multi.cu -- one host thread driving every GPU of the box (SURVEY 8b "Threading": `sb_init(n_gpus, devs)`; the reference's only
// caller is a single process, tools/src/bin/cmd.rs:67-81).  A `sb_multi` owns one context per device, all joined to one NCCL
// communicator, and one worker thread per device.  The caller stays single-threaded: sb_multi_run(fn) executes fn(rank, ctx, user)
// on every worker concurrently -- the cell-sharded call sequence of include/scanb200.h ("every rank makes the SAME sequence of
// calls on its own cell shard") -- and returns when all ranks are done.  The per-rank library calls inside fn are the ordinary
// single-context entry points; their collectives meet across the worker threads.
#include <condition_variable>
#include <mutex>
#include <thread>

#include "common.cuh"

struct sb_multi {
    int n = 0;
    std::vector<sb_ctx *> ctxs;
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    uint64_t generation = 0;  // bumped per sb_multi_run
    sb_rank_fn fn = nullptr;
    void *user = nullptr;
    int pending = 0;
    bool stop = false;
    std::vector<int> rc;
    std::vector<std::string> err;
};

static void worker_main(sb_multi *mm, int rank) {
    uint64_t seen = 0;
    for (;;) {
        sb_rank_fn fn;
        void *user;
        {
            std::unique_lock<std::mutex> lk(mm->mu);
            mm->cv_work.wait(lk, [&] { return mm->stop || mm->generation != seen; });
            if (mm->stop) return;
            seen = mm->generation;
            fn = mm->fn;
            user = mm->user;
        }
        cudaSetDevice(mm->ctxs[rank]->device);
        const int rc = fn(rank, mm->ctxs[rank], user);
        {
            std::lock_guard<std::mutex> lk(mm->mu);
            mm->rc[rank] = rc;
            mm->err[rank] = rc == SB_OK ? "" : sb_last_error();
            if (--mm->pending == 0) mm->cv_done.notify_all();
        }
    }
}

extern "C" int sb_multi_run(sb_multi *mm, sb_rank_fn fn, void *user) {
    if (!mm || !fn) return sb_fail(SB_ERR_INVALID_ARG, "sb_multi_run: NULL argument");
    {
        std::unique_lock<std::mutex> lk(mm->mu);
        mm->fn = fn;
        mm->user = user;
        mm->pending = mm->n;
        mm->generation++;
        mm->cv_work.notify_all();
        mm->cv_done.wait(lk, [&] { return mm->pending == 0; });
    }
    for (int r = 0; r < mm->n; r++)
        if (mm->rc[r] != SB_OK) return sb_fail(mm->rc[r], "rank %d: %s", r, mm->err[r].c_str());
    return SB_OK;
}

struct CommInitArgs {
    int n;
    char id[128];
};
static int comm_init_rank(int rank, sb_ctx *ctx, void *user) {
    CommInitArgs *a = (CommInitArgs *)user;
    return sb_comm_init(ctx, a->n, rank, a->id);
}

extern "C" int sb_multi_init(int n, const int *devices, sb_multi **out) {
    if (!out || n < 1) return sb_fail(SB_ERR_INVALID_ARG, "sb_multi_init: bad arguments");
    *out = nullptr;
    std::unique_ptr<sb_multi> mm(new sb_multi());
    mm->n = n;
    mm->rc.assign(n, SB_OK);
    mm->err.assign(n, "");
    for (int r = 0; r < n; r++) {
        sb_ctx *c = nullptr;
        const int rc = sb_init(devices ? devices[r] : r, &c);
        if (rc != SB_OK) {
            for (sb_ctx *p : mm->ctxs) sb_shutdown(p);
            return rc;
        }
        mm->ctxs.push_back(c);
    }
    for (int r = 0; r < n; r++) mm->workers.emplace_back(worker_main, mm.get(), r);
    if (n > 1) {
        CommInitArgs a;
        a.n = n;
        int rc = sb_comm_unique_id(a.id);
        if (rc == SB_OK) rc = sb_multi_run(mm.get(), comm_init_rank, &a);  // ncclCommInitRank blocks until every rank has joined: one thread each
        if (rc != SB_OK) {
            sb_multi_shutdown(mm.release());
            return rc;
        }
    }
    *out = mm.release();
    return SB_OK;
}

extern "C" int sb_multi_size(const sb_multi *mm) { return mm ? mm->n : 0; }

extern "C" int sb_multi_ctx(sb_multi *mm, int rank, sb_ctx **out) {
    if (!mm || !out || rank < 0 || rank >= mm->n) return sb_fail(SB_ERR_INVALID_ARG, "sb_multi_ctx: bad arguments");
    *out = mm->ctxs[rank];
    return SB_OK;
}

extern "C" void sb_multi_shutdown(sb_multi *mm) {
    if (!mm) return;
    {
        std::lock_guard<std::mutex> lk(mm->mu);
        mm->stop = true;
        mm->cv_work.notify_all();
    }
    for (auto &t : mm->workers)
        if (t.joinable()) t.join();
    for (sb_ctx *c : mm->ctxs) sb_shutdown(c);
    delete mm;
}
