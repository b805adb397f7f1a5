// Diagnostic: host->device copy bandwidth of pinned buffers by size, by how they were allocated / filled, and where
// their pages live (NUMA node via move_pages).  Written to find out why bench.py's 35 MB endpoint buffer copied at
// 18 GB/s while the 17.7 MB range buffer reached 54 GB/s.
// Build: nvcc -O2 -o bin/pinned_bw pinned_bw.cu     Run: bin/pinned_bw [gpu]
#include <cuda_runtime.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

static void node_histogram(void* p, size_t bytes, char* out, size_t outn) {
  const size_t page = 4096, n = (bytes + page - 1) / page;
  std::vector<void*> pages(n);
  std::vector<int> status(n, -1);
  for (size_t i = 0; i < n; ++i) pages[i] = (char*)p + i * page;
  long rc = syscall(SYS_move_pages, 0, (unsigned long)n, pages.data(), nullptr, status.data(), 0);
  int cnt[16] = {0}, other = 0;
  for (size_t i = 0; i < n; ++i) (status[i] >= 0 && status[i] < 16) ? cnt[status[i]]++ : other++;
  size_t o = snprintf(out, outn, "rc=%ld nodes:", rc);
  for (int k = 0; k < 16; ++k) if (cnt[k]) o += snprintf(out + o, outn - o, " n%d=%d", k, cnt[k]);
  if (other) snprintf(out + o, outn - o, " other=%d(e.g. %d)", other, status[0]);
}

static double h2d(void* d, const void* h, size_t bytes, cudaStream_t st, int reps = 10) {
  for (int i = 0; i < 2; ++i) CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < reps; ++i) CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));
  double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / reps;
  return bytes / s / 1e9;
}

static int gpu_numa_node(int dev) {
  char bdf[32];
  CK(cudaDeviceGetPCIBusId(bdf, sizeof(bdf), dev));
  for (char* c = bdf; *c; ++c) *c = tolower(*c);
  std::ifstream f(std::string("/sys/bus/pci/devices/") + bdf + "/numa_node");
  int node = -1;
  f >> node;
  return node;
}

static bool bind_to_node(int node) {
  std::ifstream f("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist");
  std::string s;
  if (!std::getline(f, s)) return false;
  cpu_set_t set;
  CPU_ZERO(&set);
  size_t i = 0;
  while (i < s.size()) {
    int a = strtol(s.c_str() + i, nullptr, 10), b = a;
    while (i < s.size() && isdigit(s[i])) ++i;
    if (i < s.size() && s[i] == '-') { ++i; b = strtol(s.c_str() + i, nullptr, 10); while (i < s.size() && isdigit(s[i])) ++i; }
    for (int c = a; c <= b; ++c) CPU_SET(c, &set);
    if (i < s.size() && s[i] == ',') ++i;
  }
  return sched_setaffinity(0, sizeof(set), &set) == 0;
}

int main(int argc, char** argv) {
  int dev = argc > 1 ? atoi(argv[1]) : 0;
  CK(cudaSetDevice(dev));
  cudaStream_t st;
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  int node = gpu_numa_node(dev);
  printf("gpu %d numa node %d, cpu now %d, online cpus %ld\n", dev, node, sched_getcpu(), sysconf(_SC_NPROCESSORS_ONLN));
  void* d;
  CK(cudaMalloc(&d, 160u << 20));
  const double sizes_mb[] = {8.0, 17.7, 35.4, 70.8, 141.6};
  char nh[256];
  std::vector<char> src(160u << 20);
  for (size_t i = 0; i < src.size(); i += 4096) src[i] = (char)i;
  for (int bound = 0; bound < 2; ++bound) {
    if (bound) {
      if (node < 0 || !bind_to_node(node)) { printf("cannot bind to node %d\n", node); break; }
      printf("=== process bound to the cpus of node %d (cpu now %d)\n", node, sched_getcpu());
    } else {
      printf("=== unbound\n");
    }
    for (double mb : sizes_mb) {
      size_t bytes = (size_t)(mb * 1e6) & ~(size_t)255;
      void* h;
      CK(cudaHostAlloc(&h, bytes, cudaHostAllocDefault));
      node_histogram(h, bytes, nh, sizeof(nh));
      double a = h2d(d, h, bytes, st);
      memset(h, 1, bytes);
      double b = h2d(d, h, bytes, st);
      memcpy(h, src.data(), bytes);
      double c = h2d(d, h, bytes, st);
      // fresh data every step, like a real pipeline: rewrite then copy, only the copy timed
      double tsum = 0;
      for (int i = 0; i < 5; ++i) {
        memcpy(h, src.data() + 64 * i, bytes);
        auto t0 = std::chrono::steady_clock::now();
        CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
        tsum += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      }
      printf("cudaHostAlloc %6.1f MB: untouched %5.1f | after memset %5.1f | after memcpy %5.1f | rewrite-then-copy %5.1f GB/s | %s\n",
             mb, a, b, c, bytes / (tsum / 5) / 1e9, nh);
      // the same copy cut into sub-copies
      for (size_t chunk : {(size_t)4 << 20, (size_t)8 << 20}) {
        for (int i = 0; i < 2; ++i) for (size_t o = 0; o < bytes; o += chunk) CK(cudaMemcpyAsync((char*)d + o, (char*)h + o, std::min(chunk, bytes - o), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < 10; ++i) for (size_t o = 0; o < bytes; o += chunk) CK(cudaMemcpyAsync((char*)d + o, (char*)h + o, std::min(chunk, bytes - o), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / 10;
        printf("    in %zu MB sub-copies: %5.1f GB/s\n", chunk >> 20, bytes / s / 1e9);
      }
      CK(cudaFreeHost(h));
      if (mb > 17 && mb < 72) {
        CK(cudaHostAlloc(&h, bytes, cudaHostAllocWriteCombined));
        memcpy(h, src.data(), bytes);
        printf("    write-combined: %5.1f GB/s\n", h2d(d, h, bytes, st));
        CK(cudaFreeHost(h));
        // malloc'ed (THP-advised) memory registered afterwards
        void* m = mmap(nullptr, bytes + (2u << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        void* m2 = (void*)(((uintptr_t)m + (2u << 20) - 1) & ~(uintptr_t)((2u << 20) - 1));
        madvise(m2, bytes, MADV_HUGEPAGE);
        memcpy(m2, src.data(), bytes);
        cudaError_t e = cudaHostRegister(m2, bytes, cudaHostRegisterDefault);
        if (e == cudaSuccess) {
          node_histogram(m2, bytes, nh, sizeof(nh));
          printf("    mmap+MADV_HUGEPAGE+cudaHostRegister: %5.1f GB/s | %s\n", h2d(d, m2, bytes, st), nh);
          cudaHostUnregister(m2);
        } else {
          printf("    cudaHostRegister failed: %s\n", cudaGetErrorString(e));
          cudaGetLastError();
        }
        munmap(m, bytes + (2u << 20));
      }
    }
    // two buffers alive at once (the bench holds ranges + endpoints + outputs)
    void *h1, *h2;
    size_t b1 = (size_t)17.7e6, b2 = (size_t)35.4e6;
    CK(cudaHostAlloc(&h1, b1, 0));
    CK(cudaHostAlloc(&h2, b2, 0));
    memcpy(h1, src.data(), b1);
    memcpy(h2, src.data(), b2);
    printf("two alive: 17.7 MB %5.1f GB/s, 35.4 MB %5.1f GB/s\n", h2d(d, h1, b1, st), h2d(d, h2, b2, st));
    // D2H for completeness
    {
      CK(cudaMemcpyAsync(h2, d, b2, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      auto t0 = std::chrono::steady_clock::now();
      for (int i = 0; i < 10; ++i) CK(cudaMemcpyAsync(h2, d, b2, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      printf("D2H 35.4 MB: %5.1f GB/s\n", b2 / (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / 10) / 1e9);
    }
    CK(cudaFreeHost(h1));
    CK(cudaFreeHost(h2));
  }
  return 0;
}
