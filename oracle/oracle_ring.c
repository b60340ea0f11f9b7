/* TEST INFRASTRUCTURE ONLY -- see oracle.h. Restates, single-threaded, the state machine of
 * the reference's SharedBuffer (include/freeimpala/data_structures.h:191-307) and this
 * build's trajectory record layout. Where the reference would block on a condition
 * variable, these functions return -1 so a test can assert "this call blocks". */
#include "oracle.h"

#include <stdlib.h>
#include <string.h>

struct orc_ring {
    unsigned char* slots; /* capacity * slot_bytes, zero-initialised (BufferEntry ctor, :164) */
    size_t slot_bytes;    /* entry_size * ELEMENT_SIZE (:206) */
    size_t capacity;
    size_t write_index, read_index, count; /* :196-198 */
    int draining;                          /* :201 */
};

orc_ring* orc_ring_create(size_t entry_size, size_t capacity) {
    orc_ring* r = (orc_ring*)calloc(1, sizeof(*r));
    r->slot_bytes = entry_size * ORC_ELEMENT_SIZE;
    r->capacity = capacity;
    r->slots = (unsigned char*)calloc(capacity ? capacity : 1, r->slot_bytes ? r->slot_bytes : 1);
    return r;
}

void orc_ring_destroy(orc_ring* r) {
    if (!r) return;
    free(r->slots);
    free(r);
}

int orc_ring_write(orc_ring* r, const void* data, size_t n) {
    if (r->count >= r->capacity) return -1; /* not_full.wait, :223 (no draining check there) */
    if (n > r->slot_bytes) return 0;        /* :226 / :240: too large -> false, no state change */
    memcpy(r->slots + r->write_index * r->slot_bytes, data, n); /* tail [n,slot) stays stale */
    r->write_index = (r->write_index + 1) % r->capacity;
    r->count++;
    return 1;
}

long orc_ring_read_batch(orc_ring* r, size_t m, void* out) {
    if (r->count < m) {
        if (r->draining) return 0; /* :278-280 */
        return -1;                 /* not_empty.wait, :273-275 */
    }
    unsigned char* o = (unsigned char*)out;
    for (size_t i = 0; i < m; i++) { /* :286-293: full-slot copy, FIFO, wraparound */
        memcpy(o + i * r->slot_bytes, r->slots + r->read_index * r->slot_bytes, r->slot_bytes);
        r->read_index = (r->read_index + 1) % r->capacity;
        r->count--;
    }
    return (long)m;
}

void orc_ring_set_draining(orc_ring* r) { r->draining = 1; }
size_t orc_ring_filled_count(const orc_ring* r) { return r->count; }
size_t orc_ring_slot_bytes(const orc_ring* r) { return r->slot_bytes; }

/* ---- record layout -------------------------------------------------------------------- */
void orc_decode_farmer(const void* batch, size_t m, size_t s, float* z, float* x, float* target) {
    const float* w = (const float*)batch;
    for (size_t b = 0; b < m; b++) {
        const float* slot = w + b * s * ORC_REC_WORDS;
        for (size_t t = 0; t < s; t++)
            memcpy(z + (b * s + t) * ORC_Z_DIM, slot + t * ORC_REC_WORDS, sizeof(float) * ORC_Z_DIM);
        for (size_t j = 0; j < ORC_X_DIM; j++) {
            size_t rec = j / ORC_X_PER_REC, off = j % ORC_X_PER_REC;
            x[b * ORC_X_DIM + j] = rec < s ? slot[rec * ORC_REC_WORDS + ORC_W_X + off] : 0.0f;
        }
        target[b] = slot[ORC_W_AUX];
    }
}

void orc_decode_vtrace(const void* batch, size_t m, size_t s, float* obs, float* mu_logits,
                       int32_t* action, float* reward, float* discount, float* bootstrap) {
    const float* w = (const float*)batch;
    for (size_t b = 0; b < m; b++) {
        const float* slot = w + b * s * ORC_REC_WORDS;
        for (size_t t = 0; t < s; t++) {
            const float* rec = slot + t * ORC_REC_WORDS;
            memcpy(obs + (b * s + t) * ORC_Z_DIM, rec, sizeof(float) * ORC_Z_DIM);
            memcpy(mu_logits + (b * s + t) * ORC_NUM_ACTIONS, rec + ORC_W_MU,
                   sizeof(float) * ORC_NUM_ACTIONS);
            memcpy(&action[b * s + t], rec + ORC_W_ACTION, sizeof(int32_t));
            reward[b * s + t] = rec[ORC_W_REWARD];
            discount[b * s + t] = rec[ORC_W_DISCOUNT];
        }
        bootstrap[b] = slot[(s - 1) * ORC_REC_WORDS + ORC_W_AUX];
    }
}
