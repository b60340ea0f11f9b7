// Trajectory ring: replaces SharedBuffer (reference include/freeimpala/data_structures.h:191-307).
//
// Host side: the same bounded MPMC FIFO state machine (write_index / read_index / count,
// one mutex, not_full / not_empty condition variables, draining flag), but the slot storage is
// pinned host memory mirrored by `capacity` slots in HBM. A writer reserves a slot under the
// lock, copies its bytes into the pinned slot OUTSIDE the lock (the reference copies 100 KiB
// under the mutex, data_structures.h:226-227, serialising every actor), and commits; commits
// are published in reservation order; runs of consecutive committed slots go to HBM with ONE
// cudaMemcpyAsync per run on the ring's side stream (a copy per slot put ~5 us of driver calls per
// trajectory under the ring lock: 1024 writes per step cost more than the learner step itself). readBatch (:267-300) becomes one sm_100a kernel that gathers M
// consecutive HBM slots (FIFO, wraparound) into a contiguous [M, slot_bytes] batch.
#include <condition_variable>
#include <mutex>
#include <new>
#include <vector>

#include "fi_common.cuh"

namespace fi {

// dst[j] = ring[(first_vec + j) mod ring_vecs], in 16-byte units. Consecutive slots are
// contiguous in HBM, so the FIFO-with-wraparound gather is a copy of at most two contiguous
// segments: every warp access is a fully coalesced 512-byte request on both sides.
// Algorithmic traffic: 2 * m * slot_bytes (read + write); HBM-bound.
constexpr int kGatherThreads = 256;
constexpr int kGatherUnroll = 8;  // 8 x 16 B in flight per thread

__global__ void __launch_bounds__(kGatherThreads)
gather_slots_kernel(const int4* __restrict__ ring, int4* __restrict__ dst, size_t ring_vecs,
                    size_t first_vec, size_t n_vecs) {
    const size_t stride = (size_t)gridDim.x * kGatherThreads;
    size_t j = (size_t)blockIdx.x * kGatherThreads + threadIdx.x;
    // main loop: kGatherUnroll independent loads first, then the stores
    for (; j + (kGatherUnroll - 1) * stride < n_vecs; j += kGatherUnroll * stride) {
        int4 v[kGatherUnroll];
#pragma unroll
        for (int u = 0; u < kGatherUnroll; u++) {
            size_t s = first_vec + j + u * stride;
            if (s >= ring_vecs) s -= ring_vecs;
            v[u] = ld_stream16(ring + s);
        }
#pragma unroll
        for (int u = 0; u < kGatherUnroll; u++) st_stream16(dst + j + u * stride, v[u]);
    }
    for (; j < n_vecs; j += stride) {
        size_t s = first_vec + j;
        if (s >= ring_vecs) s -= ring_vecs;
        st_stream16(dst + j, ld_stream16(ring + s));
    }
}

int launch_gather(const void* ring_base, size_t capacity, size_t slot_bytes, size_t first, size_t m,
                  void* dst, cudaStream_t stream) {
    if (m == 0) return FI_OK;
    if (!ring_base || !dst || capacity == 0 || first >= capacity || m > capacity)
        return set_error(FI_ERR_ARG, "gather: bad arguments (capacity=%zu first=%zu m=%zu)", capacity, first, m);
    if (slot_bytes % 16 != 0 || ((uintptr_t)ring_base & 15) || ((uintptr_t)dst & 15))
        return set_error(FI_ERR_ARG, "gather: slot_bytes and pointers must be 16-byte aligned");
    const size_t slot_vecs = slot_bytes / 16, n_vecs = m * slot_vecs;
    // grid: whole waves of 148 SMs x 8 resident CTAs, capped by the work available
    size_t want = (n_vecs + (size_t)kGatherThreads * kGatherUnroll - 1) / ((size_t)kGatherThreads * kGatherUnroll);
    size_t grid = want < (size_t)kNumSMs * 8 ? (want ? want : 1) : (size_t)kNumSMs * 8;
    LaunchScope ls("gather_slots_kernel", stream, 2.0 * (double)m * (double)slot_bytes, kWorkBytes);
    gather_slots_kernel<<<(unsigned)grid, kGatherThreads, 0, stream>>>(
        (const int4*)ring_base, (int4*)dst, capacity * slot_vecs, first * slot_vecs, n_vecs);
    return ls.done();
}

}  // namespace fi

// ------------------------------------------------------------------------------------------
struct fi_ring {
    int device = 0;
    size_t slot_bytes = 0, capacity = 0;
    unsigned char* host_slots = nullptr;  // pinned, capacity * slot_bytes
    unsigned char* dev_slots = nullptr;   // HBM,    capacity * slot_bytes
    unsigned char* dev_batch = nullptr;   // HBM, batch_cap * slot_bytes (grown on demand)
    size_t batch_cap = 0;
    cudaStream_t side = nullptr;     // H2D copies
    cudaStream_t learner = nullptr;  // default stream for the gather
    std::vector<cudaEvent_t> h2d_done;  // per slot: recorded after the H2D run that ENDS at this slot
    std::vector<size_t> h2d_ref;        // per slot: the slot whose event covers this slot's last H2D
    cudaEvent_t h2d_tail = nullptr;     // recorded after the most recent H2D run (covers all earlier ones)
    size_t copy_index = 0, uncopied = 0;  // published slots [copy_index, copy_index + uncopied) are not in HBM yet
    std::vector<unsigned char> committed;  // per slot: writer finished filling the pinned slot
    std::vector<size_t> commit_bytes;
    // Gathers read HBM slots that a later H2D copy will overwrite. Each gather gets a sequence number and an event (a
    // small ring of events); every slot remembers the last gather that read it, and the side stream waits for a gather
    // only before a copy run that overwrites slots that gather read. (One global "wait for the last gather" gate made
    // the copies of batches s+1.. wait until the learner stream reached gather(s), i.e. for the end of step s-1: the
    // copy engine then had only one step time per batch and stood idle the rest.)
    static constexpr int kGatherEvents = 8;
    cudaEvent_t gather_ev[kGatherEvents] = {};
    uint64_t gather_seq = 0;              // gathers issued so far
    uint64_t side_waited_seq = 0;         // the side stream is already ordered behind gathers <= this
    std::vector<uint64_t> slot_gather_seq;  // per slot: sequence number of the last gather that read it (0: none)

    std::mutex mu;
    std::condition_variable not_full, not_empty;
    size_t write_index = 0, read_index = 0, commit_index = 0;
    size_t count = 0;     // committed, readable entries (the reference's `count`)
    size_t reserved = 0;  // reserved but not yet committed
    uint64_t ticket_next = 0, consumed_total = 0;
    bool draining = false;
};

using fi::set_error;

extern "C" {

fi_ring* fi_ring_create(int device, size_t entry_size, size_t capacity) {
    if (capacity == 0 || entry_size == 0) {
        set_error(FI_ERR_ARG, "fi_ring_create: entry_size and capacity must be > 0");
        return nullptr;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error(FI_ERR_CUDA, "fi_ring_create: no CUDA device (there is no CPU fallback)");
        return nullptr;
    }
    FI_CUDA_OK_NULL(cudaSetDevice(device));
    fi_ring* r = new (std::nothrow) fi_ring();
    if (!r) return nullptr;
    r->device = device;
    r->slot_bytes = entry_size * FI_ELEMENT_SIZE;
    r->capacity = capacity;
    const size_t total = r->slot_bytes * capacity;
    auto fail = [&](const char* what, cudaError_t e) {
        set_error(FI_ERR_CUDA, "fi_ring_create: %s failed: %s", what, cudaGetErrorString(e));
        fi_ring_destroy(r);
        return (fi_ring*)nullptr;
    };
    cudaError_t e;
    if ((e = cudaHostAlloc((void**)&r->host_slots, total, cudaHostAllocPortable)) != cudaSuccess) return fail("cudaHostAlloc", e);
    memset(r->host_slots, 0, total);  // BufferEntry ctor zero-fills (data_structures.h:164)
    if ((e = cudaMalloc((void**)&r->dev_slots, total)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMemset(r->dev_slots, 0, total)) != cudaSuccess) return fail("cudaMemset", e);
    if ((e = cudaStreamCreateWithFlags(&r->side, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
    if ((e = cudaStreamCreateWithFlags(&r->learner, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
    r->h2d_done.assign(capacity, nullptr);
    for (size_t i = 0; i < capacity; i++)
        if ((e = cudaEventCreateWithFlags(&r->h2d_done[i], cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
    for (cudaEvent_t& ev : r->gather_ev)
        if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
    r->slot_gather_seq.assign(capacity, 0);
    if ((e = cudaEventCreateWithFlags(&r->h2d_tail, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
    r->h2d_ref.resize(capacity);
    for (size_t i = 0; i < capacity; i++) r->h2d_ref[i] = i;
    r->committed.assign(capacity, 0);
    r->commit_bytes.assign(capacity, 0);
    return r;
}

void fi_ring_destroy(fi_ring* r) {
    if (!r) return;
    cudaSetDevice(r->device);
    if (r->side) cudaStreamSynchronize(r->side);
    if (r->learner) cudaStreamSynchronize(r->learner);
    for (auto ev : r->h2d_done)
        if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : r->gather_ev)
        if (ev) cudaEventDestroy(ev);
    if (r->h2d_tail) cudaEventDestroy(r->h2d_tail);
    if (r->side) cudaStreamDestroy(r->side);
    if (r->learner) cudaStreamDestroy(r->learner);
    if (r->dev_batch) cudaFree(r->dev_batch);
    if (r->dev_slots) cudaFree(r->dev_slots);
    if (r->host_slots) cudaFreeHost(r->host_slots);
    delete r;
}

constexpr size_t kH2DRunSlots = 32;                 // copy once this many published slots are waiting ...
constexpr size_t kH2DRunBytes = (size_t)4 << 20;    // ... or this many bytes, whichever comes first

// Copy published-but-uncopied slots to HBM: one cudaMemcpyAsync per run of consecutive full slots (a short
// write copies only its own bytes, data_structures.h:226-227). Caller holds r->mu.
static void ring_flush_locked(fi_ring* r, bool force) {
    if (r->uncopied == 0) return;
    if (!force && r->uncopied < kH2DRunSlots && r->uncopied * r->slot_bytes < kH2DRunBytes) return;
    cudaError_t e = cudaSuccess;
    while (r->uncopied > 0 && e == cudaSuccess) {
        const size_t first = r->copy_index;
        size_t len = 0, bytes = 0;
        // a run: consecutive slots up to the ring end; a partially written slot ends the run after itself
        while (len < r->uncopied && first + len < r->capacity) {
            const size_t n = r->commit_bytes[first + len];
            len++;
            if (n != r->slot_bytes) { bytes = n; break; }
            bytes = 0;
        }
        const size_t full = (bytes == 0 && r->commit_bytes[first + len - 1] == r->slot_bytes) ? len : len - 1;
        // the HBM slots of this run may still be read by a gather that has not run yet
        uint64_t need = 0;
        for (size_t i = first; i < first + len; i++)
            if (r->slot_gather_seq[i] > need) need = r->slot_gather_seq[i];
        if (need > r->side_waited_seq) {
            // an event slot re-recorded by a newer gather only makes this wait longer (same stream order), never unsafe
            e = cudaStreamWaitEvent(r->side, r->gather_ev[need % fi_ring::kGatherEvents], 0);
            r->side_waited_seq = need;
        }
        if (e != cudaSuccess) break;
        if (full > 0)
            e = cudaMemcpyAsync(r->dev_slots + first * r->slot_bytes, r->host_slots + first * r->slot_bytes,
                                full * r->slot_bytes, cudaMemcpyHostToDevice, r->side);
        if (e == cudaSuccess && full < len && bytes > 0)
            e = cudaMemcpyAsync(r->dev_slots + (first + full) * r->slot_bytes, r->host_slots + (first + full) * r->slot_bytes,
                                bytes, cudaMemcpyHostToDevice, r->side);
        const size_t last = first + len - 1;
        if (e == cudaSuccess) e = cudaEventRecord(r->h2d_done[last], r->side);
        for (size_t i = first; i <= last; i++) r->h2d_ref[i] = last;
        r->copy_index = (last + 1) % r->capacity;
        r->uncopied -= len;
    }
    if (e == cudaSuccess) e = cudaEventRecord(r->h2d_tail, r->side);
    if (e != cudaSuccess) set_error(FI_ERR_CUDA, "fi_ring_write: H2D failed: %s", cudaGetErrorString(e));
}

// Publish committed slots in reservation order. Caller holds r->mu.
static int ring_publish_locked(fi_ring* r) {
    int published = 0;
    while (r->reserved > 0 && r->committed[r->commit_index]) {
        const size_t i = r->commit_index;
        r->committed[i] = 0;
        r->commit_index = (i + 1) % r->capacity;
        r->reserved--;
        r->count++;
        r->uncopied++;
        published++;
    }
    if (published) ring_flush_locked(r, false);
    return published;
}

// Block until the previous occupant of a pinned slot has reached HBM (the slot was consumed by readBatch before it
// could be reserved again, and readBatch flushes every pending copy, so the covering event has been recorded).
static void ring_wait_slot_copied(fi_ring* r, size_t slot) {
    cudaEvent_t ev;
    {
        std::lock_guard<std::mutex> lock(r->mu);
        ev = r->h2d_done[r->h2d_ref[slot]];
    }
    cudaEventSynchronize(ev);
}

static void* ring_reserve_locked(fi_ring* r, std::unique_lock<std::mutex>& lock, uint64_t* ticket, size_t* slot) {
    // not_full.wait(count < capacity), data_structures.h:223 -- no draining check, as in the reference
    r->not_full.wait(lock, [r] { return r->count + r->reserved < r->capacity; });
    const size_t i = r->write_index;
    r->write_index = (i + 1) % r->capacity;
    r->reserved++;
    if (ticket) *ticket = r->ticket_next;
    r->ticket_next++;
    *slot = i;
    return r->host_slots + i * r->slot_bytes;
}

static int ring_commit(fi_ring* r, size_t slot, size_t n) {
    int published;
    {
        std::lock_guard<std::mutex> lock(r->mu);
        r->committed[slot] = 1;
        r->commit_bytes[slot] = n;
        cudaSetDevice(r->device);
        published = ring_publish_locked(r);
    }
    if (published == 1) r->not_empty.notify_one();
    else if (published > 1) r->not_empty.notify_all();
    return 1;
}

static int ring_write_impl(fi_ring* r, const void* src, size_t n, bool blocking) {
    if (!r) return 0;
    size_t slot;
    {
        std::unique_lock<std::mutex> lock(r->mu, std::defer_lock);
        if (blocking) lock.lock();
        else if (!lock.try_lock() || r->count + r->reserved >= r->capacity) return 0;  // :245-249
        if (blocking) r->not_full.wait(lock, [r] { return r->count + r->reserved < r->capacity; });
        if (n > r->slot_bytes) return 0;  // :226 / :240: too large -> false, no state change
        ring_reserve_locked(r, lock, nullptr, &slot);
    }
    // The pinned slot may still be the source of an in-flight H2D from its previous occupant.
    cudaSetDevice(r->device);
    ring_wait_slot_copied(r, slot);
    if (n) memcpy(r->host_slots + slot * r->slot_bytes, src, n);  // bytes [n, slot) keep old content
    return ring_commit(r, slot, n);
}

int fi_ring_write(fi_ring* ring, const void* src, size_t n) { return ring_write_impl(ring, src, n, true); }
int fi_ring_try_write(fi_ring* ring, const void* src, size_t n) { return ring_write_impl(ring, src, n, false); }

size_t fi_ring_write_many(fi_ring* ring, const void* src, size_t count, size_t stride, size_t n) {
    size_t done = 0;
    for (; done < count; done++)
        if (!ring_write_impl(ring, static_cast<const unsigned char*>(src) + done * stride, n, true)) break;
    return done;
}

void* fi_ring_reserve(fi_ring* r, uint64_t* ticket) {
    if (!r) return nullptr;
    size_t slot;
    void* p;
    {
        std::unique_lock<std::mutex> lock(r->mu);
        p = ring_reserve_locked(r, lock, ticket, &slot);
    }
    cudaSetDevice(r->device);
    ring_wait_slot_copied(r, slot);
    return p;
}

int fi_ring_commit(fi_ring* r, uint64_t ticket, size_t n) {
    if (!r || n > r->slot_bytes) return 0;
    return ring_commit(r, (size_t)(ticket % r->capacity), n);
}

size_t fi_ring_reserve_many(fi_ring* r, size_t count, void** slots, uint64_t* first_ticket) {
    if (!r || count == 0 || count > r->capacity) return 0;
    size_t first;
    {
        std::unique_lock<std::mutex> lock(r->mu);
        r->not_full.wait(lock, [&] { return r->count + r->reserved + count <= r->capacity; });
        first = r->write_index;
        r->write_index = (first + count) % r->capacity;
        r->reserved += count;
        if (first_ticket) *first_ticket = r->ticket_next;
        r->ticket_next += count;
    }
    cudaSetDevice(r->device);
    for (size_t i = 0; i < count; i++) {
        const size_t slot = (first + i) % r->capacity;
        ring_wait_slot_copied(r, slot);  // the previous occupant has reached HBM
        if (slots) slots[i] = r->host_slots + slot * r->slot_bytes;
    }
    return count;
}

int fi_ring_commit_many(fi_ring* r, uint64_t first_ticket, size_t count, size_t n) {
    if (!r || n > r->slot_bytes || count == 0 || count > r->capacity) return 0;
    int published;
    {
        std::lock_guard<std::mutex> lock(r->mu);
        for (size_t i = 0; i < count; i++) {
            const size_t slot = (size_t)((first_ticket + i) % r->capacity);
            r->committed[slot] = 1;
            r->commit_bytes[slot] = n;
        }
        cudaSetDevice(r->device);
        published = ring_publish_locked(r);
    }
    if (published == 1) r->not_empty.notify_one();
    else if (published > 1) r->not_empty.notify_all();
    return 1;
}

int fi_ring_read_batch(fi_ring* r, size_t batch_size, void* stream, fi_batch* out) {
    if (!r || !out) return set_error(FI_ERR_ARG, "fi_ring_read_batch: null argument");
    memset(out, 0, sizeof(*out));
    out->slot_bytes = r->slot_bytes;
    if (batch_size == 0 || batch_size > r->capacity)
        return set_error(FI_ERR_ARG, "fi_ring_read_batch: batch_size %zu not in [1, capacity=%zu]", batch_size, r->capacity);
    cudaStream_t st = stream ? (cudaStream_t)stream : r->learner;
    out->stream = st;
    std::unique_lock<std::mutex> lock(r->mu);
    r->not_empty.wait(lock, [&] { return r->count >= batch_size || r->draining; });  // :273-275
    if (r->draining && r->count < batch_size) return 0;                              // :278-280
    FI_CUDA_OK(cudaSetDevice(r->device));
    if (r->batch_cap < batch_size) {  // first call (or a larger M): (re)allocate the batch buffer
        if (r->dev_batch) {
            FI_CUDA_OK(cudaStreamSynchronize(st));
            FI_CUDA_OK(cudaFree(r->dev_batch));
            r->dev_batch = nullptr;
            r->batch_cap = 0;
        }
        FI_CUDA_OK(cudaMalloc((void**)&r->dev_batch, batch_size * r->slot_bytes));
        r->batch_cap = batch_size;
    }
    const size_t first = r->read_index;
    // every published slot goes to HBM now. Copies are issued in FIFO order on one stream, so the event of the run that
    // holds this batch's LAST slot covers the whole batch; waiting for the tail instead would also wait for the slots
    // producers have already committed for later batches (they run up to capacity - M slots ahead), i.e. stall the
    // learner behind host->device traffic it does not need yet
    ring_flush_locked(r, true);
    const size_t last_slot = (first + batch_size - 1) % r->capacity;
    FI_CUDA_OK(cudaStreamWaitEvent(st, r->h2d_done[r->h2d_ref[last_slot]], 0));
    FI_TRY(fi::launch_gather(r->dev_slots, r->capacity, r->slot_bytes, first, batch_size, r->dev_batch, st));
    const uint64_t seq = ++r->gather_seq;
    FI_CUDA_OK(cudaEventRecord(r->gather_ev[seq % fi_ring::kGatherEvents], st));
    for (size_t i = 0; i < batch_size; i++) r->slot_gather_seq[(first + i) % r->capacity] = seq;
    r->read_index = (first + batch_size) % r->capacity;
    r->count -= batch_size;
    out->dev_ptr = r->dev_batch;
    out->num_slots = batch_size;
    out->seq = r->consumed_total;
    r->consumed_total += batch_size;
    lock.unlock();
    r->not_full.notify_all();  // :296-297
    return 1;
}

void fi_ring_set_draining(fi_ring* r) {
    if (!r) return;
    {
        std::lock_guard<std::mutex> lock(r->mu);
        r->draining = true;
    }
    r->not_empty.notify_all();
    r->not_full.notify_all();
}

size_t fi_ring_filled_count(fi_ring* r) {
    if (!r) return 0;
    std::lock_guard<std::mutex> lock(r->mu);
    return r->count;
}
size_t fi_ring_slot_bytes(const fi_ring* r) { return r ? r->slot_bytes : 0; }
size_t fi_ring_capacity(const fi_ring* r) { return r ? r->capacity : 0; }

int fi_batch_to_host(const fi_batch* b, void* dst, size_t n) {
    if (!b || !dst) return set_error(FI_ERR_ARG, "fi_batch_to_host: null argument");
    if (n > b->num_slots * b->slot_bytes) return set_error(FI_ERR_ARG, "fi_batch_to_host: n exceeds the batch");
    if (n == 0) return FI_OK;
    FI_CUDA_OK(cudaMemcpyAsync(dst, b->dev_ptr, n, cudaMemcpyDeviceToHost, (cudaStream_t)b->stream));
    FI_CUDA_OK(cudaStreamSynchronize((cudaStream_t)b->stream));
    return FI_OK;
}

int fi_op_gather(const void* ring_base, size_t capacity, size_t slot_bytes, size_t first, size_t m, void* dst, void* stream) {
    return fi::launch_gather(ring_base, capacity, slot_bytes, first, m, dst, (cudaStream_t)stream);
}

}  // extern "C"
