// multigpu.cu -- the multi-GPU entry points of libsplat_b200.so (BASELINE.json north_star: "views shard
// naturally one-per-GPU ... with an NCCL broadcast of the gaussian buffer at startup only").
//
// The reference is a single-GPU, single-context C++ program: its start-up is App::Init -> GaussianCloud::ImportPly ->
// SplatRenderer::Init (src/app.cpp:535-559).  The multi-GPU form of that call site is: every rank (one process or one
// thread per GPU) creates a SplatRenderer; ONE rank imports the PLY; SplatRenderer::InitBroadcast replicates the cloud
// with a single ncclBroadcast over NVLink 5 / NVSwitch into device memory and packs it there.  Nothing is collective on
// the frame path: a view needs the whole cloud and no other view.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): the library has no link-time dependency on it, a single-GPU host
// never loads it, and inside a process that already carries NCCL (PyTorch) the same copy is used.
//
// Also here: NUMA placement of the pinned framebuffer destination.  On the 2-socket B200 box GPUs 0-3 hang off socket 0
// and GPUs 4-7 off socket 1; a pinned buffer whose pages sit on the other socket makes every framebuffer D->H copy cross
// the socket interconnect (round 1: 51 GB/s per GPU alone, 38.5 GB/s with 8 ranks).  spl_host_alloc_near() puts the pages
// on the GPU's own node, spl_bind_thread_near() moves the calling thread there.
#include <dlfcn.h>
#include <nccl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#include "kernels.h"
#include "multigpu.h"
#include "splat_b200.h"

namespace splat
{

// ---------------------------------------------------------------------------------------------------
// NCCL, bound at run time
// ---------------------------------------------------------------------------------------------------
struct NcclApi
{
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

static NcclApi& nccl()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // a copy the process already carries (PyTorch's, or one the host linked) wins; SPL_NCCL_LIBRARY names another
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names)
            if ((api.handle = dlopen(nm, RTLD_NOW | RTLD_NOLOAD))) break;
        if (!api.handle && getenv("SPL_NCCL_LIBRARY")) api.handle = dlopen(getenv("SPL_NCCL_LIBRARY"), RTLD_NOW | RTLD_LOCAL);
        for (const char* nm : names)
            if (!api.handle) api.handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        if (!api.handle)
        {
            api.error = std::string("libnccl.so.2 could not be loaded: ") + (dlerror() ? dlerror() : "?");
            return;
        }
        auto sym = [&](const char* n) {
            void* p = dlsym(api.handle, n);
            if (!p && api.error.empty()) api.error = std::string("libnccl lacks ") + n;
            return p;
        };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
        api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    });
    return api;
}

static std::string nccl_err(const char* what, ncclResult_t r)
{
    NcclApi& a = nccl();
    return std::string(what) + ": " + (a.GetErrorString ? a.GetErrorString(r) : "NCCL error") + " (" + std::to_string((int)r) + ")";
}

static_assert(sizeof(ncclUniqueId) == SPL_COMM_ID_BYTES, "SPL_COMM_ID_BYTES must equal sizeof(ncclUniqueId)");

}  // namespace splat

using splat::nccl;
using splat::nccl_err;

bool SplatComm::GetUniqueId(unsigned char id[SPL_COMM_ID_BYTES], std::string* err)
{
    auto& a = nccl();
    if (!a.error.empty()) { if (err) *err = a.error; return false; }
    ncclUniqueId u;
    const ncclResult_t r = a.GetUniqueId(&u);
    if (r != ncclSuccess) { if (err) *err = nccl_err("ncclGetUniqueId", r); return false; }
    std::memcpy(id, &u, sizeof u);
    return true;
}

SplatComm::~SplatComm() { Destroy(); }

bool SplatComm::Init(const unsigned char id[SPL_COMM_ID_BYTES], int rankIn, int nranksIn, int deviceIn)
{
    Destroy();
    auto& a = nccl();
    if (!a.error.empty()) { error = a.error; return false; }
    if (nranksIn < 1 || rankIn < 0 || rankIn >= nranksIn) { error = "bad rank / nranks"; return false; }
    cudaError_t ce = cudaSetDevice(deviceIn);
    if (ce != cudaSuccess) { error = std::string("cudaSetDevice: ") + cudaGetErrorString(ce) + " (no CPU fallback exists)"; cudaGetLastError(); return false; }
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof u);
    ncclComm_t c = nullptr;
    const ncclResult_t r = a.CommInitRank(&c, nranksIn, u, rankIn);
    if (r != ncclSuccess) { error = nccl_err("ncclCommInitRank", r); return false; }
    comm = c;
    rank = rankIn;
    nranks = nranksIn;
    device = deviceIn;
    error.clear();
    return true;
}

void SplatComm::Destroy()
{
    if (comm)
    {
        cudaSetDevice(device);
        nccl().CommDestroy(static_cast<ncclComm_t>(comm));
    }
    comm = nullptr;
}

int SplatComm::NcclVersion()
{
    auto& a = nccl();
    int v = 0;
    if (a.error.empty() && a.GetVersion) a.GetVersion(&v);
    return v;
}

// Replicate `bytes` of device memory from `root` to every rank of the communicator (one ncclBroadcast).
bool SplatComm::Broadcast(void* deviceBuf, size_t bytes, int root, void* cudaStream)
{
    if (!comm) { error = "communicator not initialised"; return false; }
    if (root < 0 || root >= nranks) { error = "bad root"; return false; }
    const ncclResult_t r = nccl().Broadcast(deviceBuf, deviceBuf, bytes, ncclUint8, root, static_cast<ncclComm_t>(comm), static_cast<cudaStream_t>(cudaStream));
    if (r != ncclSuccess) { error = nccl_err("ncclBroadcast", r); return false; }
    return true;
}

namespace splat
{
// ---------------------------------------------------------------------------------------------------
// NUMA placement (no libnuma in the image: sysfs + raw mbind / sched_setaffinity)
// ---------------------------------------------------------------------------------------------------
namespace
{
std::string slurp_line(const std::string& path)
{
    std::ifstream f(path);
    std::string s;
    if (f) std::getline(f, s);
    return s;
}

// "0-31,64-95" -> cpu list
std::vector<int> parse_cpu_list(const std::string& s)
{
    std::vector<int> out;
    std::stringstream ss(s);
    std::string part;
    while (std::getline(ss, part, ','))
    {
        int a = 0, b = 0;
        if (std::sscanf(part.c_str(), "%d-%d", &a, &b) == 2) { for (int c = a; c <= b && out.size() < 4096; c++) out.push_back(c); }
        else if (std::sscanf(part.c_str(), "%d", &a) == 1) out.push_back(a);
    }
    return out;
}

struct HostBlock
{
    size_t bytes;
    bool mapped;  // mmap + cudaHostRegister (true) or cudaHostAlloc (false)
};
std::mutex g_hostMutex;
std::map<void*, HostBlock> g_hostBlocks;
}  // namespace

int numa_node_of_device(int device)
{
    char bus[32] = {};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) { cudaGetLastError(); return -1; }
    std::string id(bus);
    for (char& c : id) c = (char)tolower((unsigned char)c);
    const std::string s = slurp_line("/sys/bus/pci/devices/" + id + "/numa_node");
    if (s.empty()) return -1;
    return std::atoi(s.c_str());  // -1 when the platform reports none
}

bool bind_thread_near(int device, int* nodeOut, int* cpusOut)
{
    const int node = numa_node_of_device(device);
    if (nodeOut) *nodeOut = node;
    if (cpusOut) *cpusOut = 0;
    if (node < 0) return false;
    const std::vector<int> cpus = parse_cpu_list(slurp_line("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist"));
    if (cpus.empty()) return false;
    // keep to the CPUs this process may use at all (cgroup / taskset)
    cpu_set_t allowed, want;
    CPU_ZERO(&allowed);
    CPU_ZERO(&want);
    if (sched_getaffinity(0, sizeof allowed, &allowed) != 0) return false;
    int n = 0;
    for (int c : cpus)
        if (c < CPU_SETSIZE && CPU_ISSET(c, &allowed)) { CPU_SET(c, &want); n++; }
    if (!n) return false;
    if (sched_setaffinity(0, sizeof want, &want) != 0) return false;
    if (cpusOut) *cpusOut = n;
    return true;
}

void* host_alloc_near(size_t bytes, int device, int* nodeOut)
{
    if (!bytes) bytes = 1;
    const int node = numa_node_of_device(device);
    if (nodeOut) *nodeOut = -1;
    void* p = nullptr;
    if (node >= 0 && node < 1024)
    {
        const size_t page = (size_t)sysconf(_SC_PAGESIZE);
        const size_t len = (bytes + page - 1) / page * page;
        void* m = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (m != MAP_FAILED)
        {
            // MPOL_BIND (2) to the GPU's node, then first-touch: the pages exist on that node before they are pinned
            unsigned long mask[16] = {};
            mask[node / (8 * sizeof(unsigned long))] = 1ul << (node % (8 * sizeof(unsigned long)));
            const long rc = syscall(SYS_mbind, m, len, 2 /* MPOL_BIND */, mask, (unsigned long)(8 * sizeof mask), 0u);
            if (rc == 0)
            {
                std::memset(m, 0, len);
                if (cudaHostRegister(m, len, cudaHostRegisterPortable) == cudaSuccess)
                {
                    p = m;
                    if (nodeOut) *nodeOut = node;
                    std::lock_guard<std::mutex> g(g_hostMutex);
                    g_hostBlocks[p] = HostBlock{len, true};
                    return p;
                }
                cudaGetLastError();
            }
            munmap(m, len);
        }
    }
    // no NUMA information (single-node box, container without sysfs / mbind): plain pinned memory
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    std::lock_guard<std::mutex> g(g_hostMutex);
    g_hostBlocks[p] = HostBlock{bytes, false};
    return p;
}

void host_free_any(void* p)
{
    if (!p) return;
    HostBlock b{0, false};
    {
        std::lock_guard<std::mutex> g(g_hostMutex);
        auto it = g_hostBlocks.find(p);
        if (it != g_hostBlocks.end()) { b = it->second; g_hostBlocks.erase(it); }
    }
    if (b.mapped)
    {
        cudaHostUnregister(p);
        munmap(p, b.bytes);
    }
    else
    {
        cudaFreeHost(p);
    }
}

}  // namespace splat
