// nn_bench.cpp -- BASELINE config 4 through the C++ mirror: feedforward 784-512-10, batch 8192, f32,
// RecordTensor reverse-mode (forward + derivatives() + SGD update + clear/reset per step), written the way the
// reference's neural_networks.rs example is (container-level ops; inputs constants, weights tracked).
//   eager : every operator call launches its kernels (the reference's op granularity: ~50 dependent launches)
//   graph : the same step recorded once with WengertList::record and replayed as one CUDA graph launch
// Both read the loss back every step (asynchronously, one step late, so the host never stalls the device) and must
// produce the same losses bit for bit.  Prints one JSON object.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

#include "easyml.hpp"

using namespace easyml;
using T = float;

static RecordTensor<T> sigmoid(const RecordTensor<T>& z) { return ((-z).exp() + T(1)).div_swapped(T(1)); }

struct Result {
    double ms_per_step, launches_per_step, first, last, blocked_ms;
};

static Result run(int batch, int steps, bool graph) {
    const int warmup = 3;
    std::mt19937 rng(0x5EED0004);
    std::uniform_real_distribution<float> u01(0.f, 1.f), uw(-0.05f, 0.05f);
    std::vector<T> x((size_t)batch * 784), lab((size_t)batch * 10, 0.f), w1(784 * 512), w2(512 * 10);
    for (auto& v : x) v = u01(rng);
    for (int i = 0; i < batch; i++) lab[(size_t)i * 10 + rng() % 10] = 1.f;
    for (auto& v : w1) v = uw(rng);
    for (auto& v : w2) v = uw(rng);
    WengertList<T> list;
    auto W1 = RecordTensor<T>::variables(list, Tensor<T>({{"i", 784}, {"h", 512}}, w1));
    auto W2 = RecordTensor<T>::variables(list, Tensor<T>({{"h", 512}, {"o", 10}}, w2));
    auto X = RecordTensor<T>::constants(Tensor<T>({{"n", batch}, {"i", 784}}, x), list);
    auto Y = RecordTensor<T>::constants(Tensor<T>({{"n", batch}, {"o", 10}}, lab), list);
    auto L = RecordTensor<T>::constants(Tensor<T>({{"u", 1}, {"n", batch}}, std::vector<T>(batch, 1.f)), list);
    auto R = RecordTensor<T>::constants(Tensor<T>({{"o", 10}, {"v", 1}}, std::vector<T>(10, 1.f)), list);
    T* slots = nullptr;  // page-locked: the loss of step i lands in slots[i & 1]
    detail::check(eml_host_alloc(reinterpret_cast<void**>(&slots), 2 * sizeof(T)));
    double first = 0, last = 0, blocked = 0;
    void* pending = nullptr;
    int pending_slot = 0, n_step = 0;
    auto collect = [&]() {
        if (!pending) return;
        auto w0 = std::chrono::steady_clock::now();
        detail::check(eml_arrival_wait(pending));
        blocked += std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count();
        last = slots[pending_slot];
        pending = nullptr;
    };
    // the step as the reference example writes it; `slot` receives the loss
    auto body = [&](int slot) -> void* {
        auto out = sigmoid(sigmoid(X * W1) * W2);
        auto diff = out - Y;
        auto loss = ((L * diff.elementwise_multiply(diff)) * R) / T(batch);
        auto d = loss.derivatives_for(0, 0);
        sgd_update(W1, d, T(0.05));
        sgd_update(W2, d, T(0.05));
        void* arrival = loss.numbers_async(slots + slot);  // D2H of the loss: the step's result
        list.clear();
        W1.reset();
        W2.reset();
        return arrival;
    };
    auto eager_step = [&]() {
        const int slot = n_step++ & 1;
        void* arrival = body(slot);
        collect();  // the PREVIOUS step's loss
        pending = arrival;
        pending_slot = slot;
    };
    for (int i = 0; i < warmup; i++) {
        eager_step();
        if (i == 0) {
            collect();
            first = last;
        }
    }
    collect();
    detail::check(eml_sync());
    Result r{};
    const long long l0 = eml_kernel_launches();
    blocked = 0;
    if (!graph) {
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < steps; i++) eager_step();
        collect();
        detail::check(eml_sync());
        r.ms_per_step = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / steps * 1e3;
    } else {
        // two recorded steps (one per read-back slot), replayed alternately
        RecordedStep g0 = list.record([&] { body(0); });
        RecordedStep g1 = list.record([&] { body(1); });
        const long long l1 = eml_kernel_launches();
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < steps; i++) {
            const int slot = i & 1;
            void* arrival = (slot ? g1 : g0).launch();
            collect();
            pending = arrival;
            pending_slot = slot;
        }
        collect();
        detail::check(eml_sync());
        r.ms_per_step = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / steps * 1e3;
        r.launches_per_step = double(eml_kernel_launches() - l1) / steps;
        if (std::getenv("NN_PROFILE")) std::fprintf(stderr, "%s", g0.profile(200).c_str());
    }
    if (!graph) r.launches_per_step = double(eml_kernel_launches() - l0) / steps;
    r.first = first;
    r.last = last;
    r.blocked_ms = blocked / steps * 1e3;
    eml_host_free(slots);
    return r;
}

int main(int argc, char** argv) {
    const int batch = argc > 1 ? std::atoi(argv[1]) : 8192, steps = argc > 2 ? std::atoi(argv[2]) : 20;
    try {
        Result e = run(batch, steps, false);
        Result g = run(batch, steps, true);
        std::printf(
            "{\"steps_per_s\": %.3f, \"ms_per_step\": %.5f, \"launches_per_step\": %.1f, \"loss_first\": %.7g, "
            "\"loss_last\": %.7g, \"host_blocked_on_device_ms_per_step\": %.5f, \"batch\": %d, \"steps\": %d, "
            "\"loss_readback\": \"every step, asynchronous, read one step later\", \"host\": \"C++ mirror easyml.hpp\", "
            "\"graph\": {\"steps_per_s\": %.3f, \"ms_per_step\": %.5f, \"launches_per_step\": %.1f, \"loss_last\": %.7g, "
            "\"same_loss_as_eager_bit_for_bit\": %s, \"what\": \"the same step recorded once (WengertList::record -> CUDA "
            "graph, two instances for the two read-back slots) and replayed\"}}\n",
            1e3 / e.ms_per_step, e.ms_per_step, e.launches_per_step, e.first, e.last, e.blocked_ms, batch, steps,
            1e3 / g.ms_per_step, g.ms_per_step, g.launches_per_step, g.last,
            std::memcmp(&e.last, &g.last, sizeof(double)) == 0 ? "true" : "false");
    } catch (const std::exception& e) {
        std::fprintf(stderr, "nn_bench: %s\n", e.what());
        return 1;
    }
    return 0;
}
