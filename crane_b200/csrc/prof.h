// Forward-pass profiling, the B200 counterpart of crane-core/src/ops/prof.rs:1-61,99-197 (`CRANE_PROF=1`, `CRANE_PROF_EVERY`).
//
// The reference times each pass twice on the host -- `enqueue` (last op submitted) and `wall` (after a device sync) -- and charges
// host wall time to 15 named spans in three tiers (embed norm attn gdn mlp resid head | proj conv qkv recur finish | prep launch
// post).  Here every span mark also records a cudaEvent on the engine stream, so the same names carry DEVICE time as well: the
// question "dispatch-bound or kernel-bound" is answered per stage, and an `ncu` launch list lines up with the stage names.  Marks
// are leaf spans (finer than the reference's: attention and MLP are split into their kernels); the reference's tiers are sums.
// An event between two kernels removes their programmatic (PDL) overlap, so a profiled pass runs a few percent slower than an
// unprofiled one: the report states both the profiled wall time and nothing else -- never quote it as a benchmark number.
#pragma once

#include <cuda_runtime.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

namespace cb {

enum SpanId : int {
    SP_EMBED = 0, SP_NORM, SP_ATTN_QKV, SP_ATTN_ROPE, SP_ATTN_FLASH, SP_ATTN_O, SP_GDN_PROJ, SP_GDN_CONV, SP_GDN_QKV, SP_GDN_RECUR,
    SP_GDN_FINISH, SP_MLP_GATE_UP, SP_MLP_DOWN, SP_SPLICE, SP_HEAD, SP_VIT_STAGE, SP_VIT_PATCH, SP_VIT_NORM, SP_VIT_QKV, SP_VIT_ROPE,
    SP_VIT_FLASH, SP_VIT_PROJ, SP_VIT_FC1, SP_VIT_FC2, SP_VIT_MERGER, SP_DECODE, SP_OTHER, SP_COUNT
};
static const char* const SPAN_NAMES[SP_COUNT] = {
    "embed", "norm", "attn.qkv", "attn.rope", "attn.flash", "attn.o", "gdn.proj", "gdn.conv", "gdn.qkv", "gdn.recur", "gdn.finish",
    "mlp.gate_up", "mlp.down", "splice", "head", "vit.stage", "vit.patch", "vit.norm", "vit.qkv", "vit.rope", "vit.flash", "vit.proj",
    "vit.fc1", "vit.fc2", "vit.merger", "decode", "other"};

// kernels outside the engine (gdn.cu) mark their own stages through this sink; it is empty unless a profiled pass is running
struct SpanSink { void (*fn)(void*, int) = nullptr; void* ctx = nullptr; };
inline SpanSink& span_sink() { static thread_local SpanSink s; return s; }
inline void span_mark(int id) { SpanSink& s = span_sink(); if (s.fn) s.fn(s.ctx, id); }

struct PassTotals {
    uint64_t passes = 0, tokens = 0;
    double enqueue_ms = 0, wall_ms = 0, device_ms = 0;
    double host_ms[SP_COUNT] = {}, dev_ms[SP_COUNT] = {};
};

class PassProfiler {
  public:
    bool on = false;
    int every = 64;
    PassTotals window[2], total[2];          // [0] decode (one position), [1] prefill -- as prof.rs `kind`

    void init_from_env() {
        const char* e = getenv("CRANE_PROF");
        on = e && e[0] && std::string(e) != "0";
        if (const char* n = getenv("CRANE_PROF_EVERY")) { const int v = atoi(n); if (v > 0) every = v; }
    }
    bool active() const { return in_pass; }

    void begin(cudaStream_t st) {
        if (!on || in_pass) return;
        stream = st; in_pass = true; n_ev = 0;
        t0 = last = clock::now();
        cur = SP_OTHER;
        for (double& v : host_acc) v = 0;
        span_sink() = SpanSink{[](void* c, int id) { static_cast<PassProfiler*>(c)->mark(id); }, this};
        push_event(SP_OTHER);
    }
    // everything enqueued from here to the next mark belongs to `id`
    void mark(int id) {
        if (!in_pass) return;
        const auto now = clock::now();
        host_acc[cur] += std::chrono::duration<double, std::milli>(now - last).count();
        last = now; cur = id;
        push_event(id);
    }
    // `tokens` positions were processed; returns false when the sync failed (sample dropped, as the reference does)
    bool end(size_t tokens, int kind_override = -1, uint64_t n_passes = 1) {
        if (!in_pass) return true;
        mark(SP_OTHER);
        in_pass = false;
        span_sink() = SpanSink{};
        const double enq = std::chrono::duration<double, std::milli>(clock::now() - t0).count();
        if (cudaStreamSynchronize(stream) != cudaSuccess) { fprintf(stderr, "[crane-prof] device sync failed, dropping sample\n"); return false; }
        const double wall = std::chrono::duration<double, std::milli>(clock::now() - t0).count();
        const int kind = kind_override >= 0 ? kind_override : (tokens > 1 ? 1 : 0);
        for (PassTotals* t : {&window[kind], &total[kind]}) {
            t->passes += n_passes; t->tokens += tokens; t->enqueue_ms += enq; t->wall_ms += wall;
            for (int i = 0; i < SP_COUNT; ++i) t->host_ms[i] += host_acc[i];
        }
        for (int i = 0; i + 1 < n_ev; ++i) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ev[i], ev[i + 1]) != cudaSuccess) continue;
            window[kind].dev_ms[span_of[i]] += ms; total[kind].dev_ms[span_of[i]] += ms;
            window[kind].device_ms += ms; total[kind].device_ms += ms;
        }
        if (window[kind].passes >= (uint64_t)every) {
            fprintf(stderr, "%s\n", line(kind, window[kind]).c_str());
            window[kind] = PassTotals{};
        }
        return true;
    }
    void release() {
        for (cudaEvent_t e : ev) cudaEventDestroy(e);
        ev.clear();
    }

    // one summary line per kind, normalised per pass -- the reference's three tiers first, the finer leaf spans after
    static std::string line(int kind, const PassTotals& t) {
        char b[256];
        const double n = t.passes ? (double)t.passes : 1.0;
        std::string s = "[crane-prof] ";
        snprintf(b, sizeof b, "%s x%llu (%.1f tok/pass): enqueue %.3f ms, wall %.3f ms, device %.3f ms |", kind ? "prefill" : "decode",
                 (unsigned long long)t.passes, (double)t.tokens / n, t.enqueue_ms / n, t.wall_ms / n, t.device_ms / n);
        s += b;
        auto grp = [&](const double* a, int lo, int hi) { double v = 0; for (int i = lo; i <= hi; ++i) v += a[i]; return v / n; };
        struct G { const char* name; int lo, hi; };
        static const G tier1[] = {{"embed", SP_EMBED, SP_EMBED}, {"norm", SP_NORM, SP_NORM}, {"attn", SP_ATTN_QKV, SP_ATTN_O},
                                  {"gdn", SP_GDN_PROJ, SP_GDN_FINISH}, {"mlp", SP_MLP_GATE_UP, SP_MLP_DOWN}, {"head", SP_HEAD, SP_HEAD},
                                  {"vit", SP_VIT_STAGE, SP_VIT_MERGER}, {"decode", SP_DECODE, SP_DECODE}};
        s += " device ms:";
        for (const G& g : tier1) { snprintf(b, sizeof b, " %s %.3f", g.name, grp(t.dev_ms, g.lo, g.hi)); s += b; }
        s += " resid 0 (fused into the o / down epilogues) | gdn: proj conv qkv recur finish =";
        for (int i = SP_GDN_PROJ; i <= SP_GDN_FINISH; ++i) { snprintf(b, sizeof b, " %.3f", t.dev_ms[i] / n); s += b; }
        s += " | recur: prep 0 launch = recur post 0 (one kernel) | leaves (device / host ms):";
        for (int i = 0; i < SP_COUNT; ++i)
            if (t.dev_ms[i] > 0 || t.host_ms[i] > 0) { snprintf(b, sizeof b, " %s %.3f/%.3f", SPAN_NAMES[i], t.dev_ms[i] / n, t.host_ms[i] / n); s += b; }
        return s;
    }
    static std::string json(const PassTotals t[2]) {
        std::string s = "{";
        char b[160];
        for (int kind = 0; kind < 2; ++kind) {
            const PassTotals& p = t[kind];
            const double n = p.passes ? (double)p.passes : 1.0;
            snprintf(b, sizeof b, "%s\"%s\": {\"passes\": %llu, \"tokens\": %llu, \"enqueue_ms\": %.4f, \"wall_ms\": %.4f, \"device_ms\": %.4f, \"spans\": {",
                     kind ? ", " : "", kind ? "prefill" : "decode", (unsigned long long)p.passes, (unsigned long long)p.tokens, p.enqueue_ms / n,
                     p.wall_ms / n, p.device_ms / n);
            s += b;
            bool first = true;
            for (int i = 0; i < SP_COUNT; ++i) {
                if (p.dev_ms[i] == 0 && p.host_ms[i] == 0) continue;
                snprintf(b, sizeof b, "%s\"%s\": {\"device_ms\": %.4f, \"host_ms\": %.4f}", first ? "" : ", ", SPAN_NAMES[i], p.dev_ms[i] / n, p.host_ms[i] / n);
                s += b; first = false;
            }
            s += "}}";
        }
        return s + "}";
    }

  private:
    using clock = std::chrono::steady_clock;
    cudaStream_t stream = nullptr;
    bool in_pass = false;
    int n_ev = 0, cur = SP_OTHER;
    clock::time_point t0, last;
    std::vector<cudaEvent_t> ev;
    std::vector<int> span_of;
    double host_acc[SP_COUNT] = {};

    void push_event(int id) {
        if (n_ev == (int)ev.size()) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) return;
            ev.push_back(e); span_of.push_back(SP_OTHER);
        }
        span_of[n_ev] = id;
        cudaEventRecord(ev[n_ev], stream);
        ++n_ev;
    }
};

}  // namespace cb
