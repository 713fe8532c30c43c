// nn_bench.cpp -- BASELINE config 4 through the C++ mirror: feedforward 784-512-10, batch 8192, f32,
// RecordTensor reverse-mode (forward + derivatives() + SGD update + clear/reset per step), written the way the
// reference's neural_networks.rs example is (container-level ops; inputs constants, weights tracked).
// Prints one JSON object.  The compiled host shows the library's step rate without interpreter overhead.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "easyml.hpp"

using namespace easyml;
using T = float;

static RecordTensor<T> sigmoid(const RecordTensor<T>& z) { return ((-z).exp() + T(1)).div_swapped(T(1)); }

int main(int argc, char** argv) {
    const int batch = argc > 1 ? std::atoi(argv[1]) : 8192, steps = argc > 2 ? std::atoi(argv[2]) : 20;
    const int warmup = 3;
    try {
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
        double first = 0, last = 0;
        // the loss of step i is copied to page-locked memory asynchronously and read once step i+1 is enqueued
        // (every step still transfers its result; the host just never stalls the device)
        T* slots = nullptr;
        detail::check(eml_host_alloc(reinterpret_cast<void**>(&slots), 2 * sizeof(T)));
        void* pending = nullptr;
        int pending_slot = 0, n_step = 0;
        const bool pipelined = argc > 3 ? std::atoi(argv[3]) != 0 : true;
        double wait_s = 0;  // time the host spent blocked on the device (loss arrival)
        auto collect = [&]() {
            if (pending) {
                auto w0 = std::chrono::steady_clock::now();
                detail::check(eml_arrival_wait(pending));
                wait_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count();
                last = slots[pending_slot];
                pending = nullptr;
            }
        };
        auto step = [&]() {
            auto out = sigmoid(sigmoid(X * W1) * W2);
            auto diff = out - Y;
            auto loss = ((L * diff.elementwise_multiply(diff)) * R) / T(batch);
            auto d = loss.derivatives_for(0, 0);
            sgd_update(W1, d, T(0.05));
            sgd_update(W2, d, T(0.05));
            if (pipelined) {
                const int slot = n_step++ & 1;
                void* arrival = loss.numbers_async(slots + slot);  // D2H of the loss: the step's result
                collect();                                         // ... of the PREVIOUS step
                pending = arrival;
                pending_slot = slot;
            } else {
                last = loss.numbers().data()[0];
            }
            list.clear();
            W1.reset();
            W2.reset();
        };
        for (int i = 0; i < warmup; i++) {
            step();
            if (i == 0) {
                collect();
                first = last;
            }
        }
        collect();
        detail::check(eml_sync());
        const long long l0 = eml_kernel_launches();
        wait_s = 0;
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < steps; i++) step();
        collect();
        detail::check(eml_sync());
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / steps;
        std::printf("{\"steps_per_s\": %.3f, \"ms_per_step\": %.5f, \"launches_per_step\": %.1f, \"loss_first\": %.7g, "
                    "\"loss_last\": %.7g, \"batch\": %d, \"steps\": %d, \"loss_readback\": \"%s\", \"host_blocked_on_device_ms_per_step\": %.5f, \"host\": \"C++ mirror easyml.hpp\"}\n",
                    1.0 / dt, dt * 1e3, double(eml_kernel_launches() - l0) / steps, first, last, batch, steps,
                    pipelined ? "every step, asynchronous, read one step later" : "every step, blocking",
                    wait_s / steps * 1e3);
        eml_host_free(slots);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "nn_bench: %s\n", e.what());
        return 1;
    }
    return 0;
}
