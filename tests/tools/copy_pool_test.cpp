// CPU unit test of the host copy pool in tray_racing_b200/csrc/host_copy.h (no CUDA call is made): random sizes and
// offsets, several caller threads at once, every copy compared with the source.
#include <cstdio>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

#include "host_copy.h"

int main() {
    std::vector<std::thread> callers;
    int bad = 0;
    for (int c = 0; c < 3; c++)
        callers.emplace_back([c, &bad] {
            std::mt19937_64 rng(1234 + c);
            std::vector<unsigned char> src(48u << 20), dst(48u << 20);
            for (auto& b : src) b = (unsigned char)rng();
            for (int it = 0; it < 40; it++) {
                const size_t len = it < 4 ? (size_t)(rng() % 4096) : (size_t)(rng() % (40u << 20)) + 1;
                const size_t so = rng() % (src.size() - len), dof = rng() % (dst.size() - len);
                memset(dst.data() + dof, 0, len);
                tray::par_copy(dst.data() + dof, src.data() + so, len);
                if (memcmp(dst.data() + dof, src.data() + so, len) != 0) { __atomic_fetch_add(&bad, 1, __ATOMIC_RELAXED); }
            }
        });
    for (auto& t : callers) t.join();
    printf(bad ? "FAILED %d\n" : "ok\n", bad);
    return bad ? 1 : 0;
}
