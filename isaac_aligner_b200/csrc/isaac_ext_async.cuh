// isaac_ext_submit_* / isaac_ext_wait: the three tile calls without blocking the caller ("async submit, wait on a
// ticket").  The reference overlaps the loading and sorting of the next tile's matches with the processing of the current one
// (SelectMatchesTransition.cpp:316-340 runs load / compute / flush slots side by side); a submitted call runs on a worker thread of
// the context (on the context's streams) while the caller's thread goes on -- e.g. with isaac_ext_prefetch_reads / _batch of the next
// tile, the only calls the context accepts from other threads meanwhile.  A context is not re-entrant (like the reference's
// per-thread builders), so one call is in flight per context: a second submit before the wait is refused.  Included by isaac_ext.cu.
#pragma once
#include <atomic>
#include <thread>

struct AsyncState
{
    std::thread worker;
    std::atomic<std::thread::id> workerId;
    std::atomic<bool> inFlight{false};
    uint64_t ticket = 0;
    int status = ISAAC_EXT_OK;
    int kind = 0;                                   // 1 build_fragments, 2 rescue_shadows, 3 build_templates
    // the small argument structs are copied at submit; the arrays they point to stay the caller's until the wait
    isaac_ext_build_batch_t batch;
    isaac_ext_tls_t tls;
    isaac_ext_template_options_t options;
    uint32_t requestCount = 0;
    const isaac_ext_rescue_request_t *requests = nullptr;
    isaac_ext_build_result_t build;
    isaac_ext_rescue_result_t rescue;
    isaac_ext_template_result_t templates;
};

/// true for every thread but the worker itself while a submitted call runs (the error text is not touched: it belongs to that call)
bool asyncCallInFlight(const isaac_ext_ctx *ctx)
{
    return ctx && ctx->async && ctx->async->inFlight && std::this_thread::get_id() != ctx->async->workerId.load();
}

void releaseAsync(AsyncState *state)
{
    if (!state) return;
    if (state->worker.joinable()) state->worker.join();
    delete state;
}

namespace
{
int submitCall(isaac_ext_ctx *ctx, int kind, uint64_t *ticketOut)
{
    AsyncState &a = *ctx->async;
    a.kind = kind;
    a.status = ISAAC_EXT_OK;
    a.workerId = std::thread::id();                 // nobody until the worker names itself
    a.inFlight = true;
    *ticketOut = ++a.ticket;
    a.worker = std::thread([ctx, kind]() {
        AsyncState &s = *ctx->async;
        s.workerId = std::this_thread::get_id();        // before its first entry point: asyncCallInFlight lets only this thread through
        s.status = kind == 1 ? isaac_ext_build_fragments(ctx, &s.batch, &s.build)
                 : kind == 2 ? isaac_ext_rescue_shadows(ctx, &s.tls, s.requestCount, s.requests, &s.rescue)
                             : isaac_ext_build_templates(ctx, &s.batch, &s.tls, &s.options, &s.templates);
    });
    return ISAAC_EXT_OK;
}

int readyToSubmit(isaac_ext_ctx *ctx, uint64_t *ticketOut)
{
    if (!ticketOut) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null ticket");
    if (!ctx->async) ctx->async = new AsyncState();
    if (ctx->async->inFlight) return ctx->fail(ISAAC_EXT_E_UNSUPPORTED, "a submitted call is in flight on this context: isaac_ext_wait first");
    return ISAAC_EXT_OK;
}
} // namespace

extern "C" int isaac_ext_submit_build_fragments(isaac_ext_ctx *ctx, const isaac_ext_build_batch_t *batch, uint64_t *ticketOut)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    if (!batch) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null batch");
    if (const int rc = readyToSubmit(ctx, ticketOut)) return rc;
    ctx->async->batch = *batch;
    return submitCall(ctx, 1, ticketOut);
}

extern "C" int isaac_ext_submit_rescue_shadows(isaac_ext_ctx *ctx, const isaac_ext_tls_t *tls, uint32_t requestCount,
                                               const isaac_ext_rescue_request_t *requests, uint64_t *ticketOut)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    if (!tls) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null template length statistics");
    if (const int rc = readyToSubmit(ctx, ticketOut)) return rc;
    ctx->async->tls = *tls; ctx->async->requestCount = requestCount; ctx->async->requests = requests;
    return submitCall(ctx, 2, ticketOut);
}

extern "C" int isaac_ext_submit_build_templates(isaac_ext_ctx *ctx, const isaac_ext_build_batch_t *batch, const isaac_ext_tls_t *tls,
                                                const isaac_ext_template_options_t *options, uint64_t *ticketOut)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    if (!batch || !tls || !options) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null argument");
    if (const int rc = readyToSubmit(ctx, ticketOut)) return rc;
    ctx->async->batch = *batch; ctx->async->tls = *tls; ctx->async->options = *options;
    return submitCall(ctx, 3, ticketOut);
}

extern "C" int isaac_ext_wait(isaac_ext_ctx *ctx, uint64_t ticket, void *resultOut)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    if (!ctx->async || !ctx->async->inFlight || ctx->async->ticket != ticket)
        return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "no submitted call with this ticket is in flight");
    AsyncState &a = *ctx->async;
    a.worker.join();
    a.inFlight = false;
    if (a.status) return a.status;                  // isaac_ext_last_error holds the text the call left
    if (!resultOut) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null result");
    if (a.kind == 1) *static_cast<isaac_ext_build_result_t *>(resultOut) = a.build;
    else if (a.kind == 2) *static_cast<isaac_ext_rescue_result_t *>(resultOut) = a.rescue;
    else *static_cast<isaac_ext_template_result_t *>(resultOut) = a.templates;
    return ISAAC_EXT_OK;
}
