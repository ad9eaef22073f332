// host_queue.cpp — the utterance queue shared by the ranks of one node (C ABI: jgpu_queue_*).
//
// BASELINE configs[3] / north_star: a batch of utterances sharded over the GPUs of a box with whole-utterance work
// stealing.  Utterances share only read-only data, so the only thing the ranks exchange while they
// decode is "which utterance is next": one 64-bit counter in a POSIX shared-memory segment, bumped with an atomic
// fetch-add by whichever rank has a lane free (jgpu_decode_queue, jgpu_engine.cu).  No data-path collective, no
// TCP round trip per claim; the collectives around it (barrier, result gather, timing reduction) stay with the
// caller's NCCL process group.
#include <atomic>
#include <cerrno>
#include <cstring>
#include <new>
#include <string>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/juicer_b200.h"
#include "jgpu_err.h"
#include "host_queue.h"

static_assert(sizeof(std::atomic<long long>) == 8 && alignof(std::atomic<long long>) <= 64, "lock-free 64-bit counter");

namespace {
int q_fail(int code, const std::string& msg)
{
    jgpu_err_buf() = msg;
    return code;
}
} // namespace

extern "C" int jgpu_queue_open(const char* name, int32_t create, jgpu_queue** out)
{
    if (!name || !out || !name[0]) return q_fail(JGPU_E_ARG, "jgpu_queue_open: bad argument");
    *out = nullptr;
    std::string shm = name[0] == '/' ? std::string(name) : "/" + std::string(name);
    for (size_t i = 1; i < shm.size(); ++i)
        if (shm[i] == '/') shm[i] = '_';
    int fd = shm_open(shm.c_str(), create ? (O_CREAT | O_RDWR) : O_RDWR, 0600);
    if (fd < 0) return q_fail(JGPU_E_IO, "shm_open(" + shm + "): " + strerror(errno));
    if (create && ftruncate(fd, sizeof(JgpuQueueShared)) != 0) {
        const std::string e = strerror(errno);
        close(fd);
        return q_fail(JGPU_E_IO, "ftruncate(" + shm + "): " + e);
    }
    struct stat st;
    if (fstat(fd, &st) != 0 || (size_t)st.st_size < sizeof(JgpuQueueShared)) {
        close(fd);
        return q_fail(JGPU_E_IO, "queue segment " + shm + " is not initialised yet (open it after the creator's barrier)");
    }
    void* p = mmap(nullptr, sizeof(JgpuQueueShared), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return q_fail(JGPU_E_IO, "mmap(" + shm + "): " + strerror(errno));
    jgpu_queue* q = new (std::nothrow) jgpu_queue;
    if (!q) { munmap(p, sizeof(JgpuQueueShared)); return q_fail(JGPU_E_ARG, "out of memory"); }
    q->shared = static_cast<JgpuQueueShared*>(p);
    q->name = shm;
    q->owner = create != 0;
    if (create) {
        q->shared->next.store(0, std::memory_order_relaxed);
        q->shared->generation.fetch_add(1, std::memory_order_release);
    }
    *out = q;
    return JGPU_OK;
}

extern "C" int jgpu_queue_reset(jgpu_queue* q)
{
    if (!q) return q_fail(JGPU_E_ARG, "null queue");
    q->shared->next.store(0, std::memory_order_release);
    q->shared->generation.fetch_add(1, std::memory_order_release);
    return JGPU_OK;
}

extern "C" int64_t jgpu_queue_claim(jgpu_queue* q, int64_t n)
{
    if (!q || n < 0) return -1;
    return (int64_t)q->shared->next.fetch_add((long long)n, std::memory_order_acq_rel);
}

extern "C" int64_t jgpu_queue_position(jgpu_queue* q)
{
    return q ? (int64_t)q->shared->next.load(std::memory_order_acquire) : -1;
}

extern "C" int jgpu_queue_close(jgpu_queue* q)
{
    if (!q) return JGPU_OK;
    munmap(q->shared, sizeof(JgpuQueueShared));
    if (q->owner) shm_unlink(q->name.c_str());
    delete q;
    return JGPU_OK;
}
