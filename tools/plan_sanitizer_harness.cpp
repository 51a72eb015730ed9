#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>
#include <cstdint>
#include "mpgpu_internal.h"
namespace mpgpu { static std::string g_err; void set_error(const std::string &m) { g_err = m; } int cuda_fail(cudaError_t, const char *) { return 3; } }
extern "C" int mpgpu_host_plan_selftest(int, const int32_t *, const int32_t *, const int32_t *, int, int, int, int, int, int);
extern "C" int mpgpu_host_visit_order(int, const int32_t *, const int32_t *, int32_t *);
// random unrooted binary tree as ring tables: stepwise insertion of tips
static void random_tree(int n, unsigned seed, std::vector<int32_t> &bn, std::vector<int32_t> &bs)
{
    mpgpu::HostTree t; t.n = n; t.bn.assign(3 * (2 * n - 1), 0); t.bs.assign(3 * (2 * n - 1), 0);
    std::mt19937 rng(seed);
    // start: tips 1,2,3 on inner node n+1
    int inner = n + 1;
    t.hookup(3 * 1, 3 * inner); t.hookup(3 * 2, 3 * inner + 1); t.hookup(3 * 3, 3 * inner + 2);
    std::vector<int> edges = {3 * 1, 3 * 2, 3 * 3};      // one ref per edge
    for (int tip = 4; tip <= n; tip++) {
        inner++;
        const int e = edges[rng() % edges.size()];
        const int f = t.back(e);
        t.hookup(e, 3 * inner); t.hookup(f, 3 * inner + 1); t.hookup(3 * tip, 3 * inner + 2);
        edges.push_back(3 * inner + 1); edges.push_back(3 * tip);
    }
    bn = t.bn; bs = t.bs;
}
static int run_all(int seed_off)
{
    for (int n : {50, 200, 600}) {
        std::vector<int32_t> bn, bs, order(2 * n - 1);
        random_tree(n, 7 + n, bn, bs);
        if (mpgpu_host_visit_order(n, bn.data(), bs.data(), order.data())) { printf("order failed\n"); return 1; }
        for (int rep = 0; rep < 20; rep++)
            for (int nt : {2, 3, 4, 8})
                for (int pieces : {1, 2, 3}) {
                    const int rc = mpgpu_host_plan_selftest(n, bn.data(), bs.data(), order.data(), 1, 2 * n - 2, 1, 6, nt, pieces);
                    if (rc) { printf("selftest rc=%d n=%d nt=%d pieces=%d: %s\n", rc, n, nt, pieces, mpgpu::g_err.c_str()); return 1; }
                }
    }
    (void)seed_off;
    return 0;
}
#include <thread>
int main()
{
    if (run_all(0)) return 1;
    // two host threads (two contexts, one per GPU) enumerating at the same time: one gets the pool, the other falls back to add()
    int rc1 = 0, rc2 = 0;
    std::thread a([&]() { rc1 = run_all(1); }), b([&]() { rc2 = run_all(2); });
    a.join(); b.join();
    if (rc1 || rc2) return 1;
    printf("ok\n");
    return 0;
}
