/*
 * rpk_oracle.c -- CPU restatement of the reference's scheduling hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under k8s-runpod-kubelet_b200/ links, loads
 * or calls this file.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may use it, and there
 * only as the checker or as the timed CPU arm, never as the product path.
 *
 * PARITY STATUS
 *   - column producers (extract_gpu_memory, validate_cloud_type,
 *     annotation fallback): PINNED by the reference's own tests
 *     (pkg/virtual_kubelet/annotations_test.go:117,141,233,121,225;
 *      runpod_test.go:89-90) -- see tests/golden/column_producers.json.
 *   - XXH64: the reference never calls a hash; the kernel's
 *     hash is pinned against the public XXH64 spec, i.e. what
 *     github.com/cespare/xxhash/v2 v2.1.2 (go.mod:60, indirect) implements,
 *     through python-xxhash 3.7.0 vectors in tests/golden/xxh64_kat.json.
 *   - selection results (gpuTypeIds) and the status changed-set:
 *     **parity unpinned** -- no reference test asserts them and the Go
 *     toolchain is absent, so the reference cannot be run here.  They are
 *     restated line by line from the cited Go source and cross-checked
 *     against an independently written numpy restatement
 *     (tests/np_restatement.py) and hand-derived KATs.
 *
 * Every function cites the reference lines it follows
 * (paths relative to /root/reference/pkg/virtual_kubelet/).
 */
#include <limits.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <ctype.h>
#include <string.h>

#define RPK_CLOUD_SECURE 0
#define RPK_CLOUD_COMMUNITY 1
#define RPK_CLOUD_OTHER 2 /* any string that is neither: runpod_client.go:469-475 leaves price=0, cloudCheck=false */

/* ------------------------------------------------------------------------ */
/* column producers                                                          */
/* ------------------------------------------------------------------------ */

/* strconv.Atoi(s) for a 64-bit Go int: optional sign, >=1 decimal digits, no
 * spaces/underscores, range error on overflow.  Returns 1 on success. */
static int go_atoi(const char *s, int64_t *out) {
    size_t n = strlen(s);
    if (n == 0) return 0;
    size_t i = 0;
    int neg = 0;
    if (s[0] == '+' || s[0] == '-') {
        neg = (s[0] == '-');
        i = 1;
        if (n == 1) return 0;
    }
    uint64_t acc = 0;
    const uint64_t lim = neg ? (uint64_t)1 << 63 : ((uint64_t)1 << 63) - 1;
    for (; i < n; ++i) {
        unsigned d = (unsigned)(s[i] - '0');
        if (d > 9) return 0;
        if (acc > (lim - d) / 10) return 0; /* ErrRange */
        acc = acc * 10 + d;
    }
    *out = neg ? (int64_t)(0 - acc) : (int64_t)acc;
    return 1;
}

/* extractGPUMemory -- runpod_client.go:1181-1191: ""->16; Atoi ok -> value
 * (may be <=0 or huge); parse error -> 16. */
int64_t rpk_oracle_extract_gpu_memory(const char *mem_str) {
    const int64_t default_memory = 16;
    if (mem_str == NULL || mem_str[0] == '\0') return default_memory;
    int64_t v;
    if (go_atoi(mem_str, &v)) return v;
    return default_memory;
}

/* The engine's pod column is int32 ("clamp to int32 on
 * ingest"); the C-ABI rejects offer tables holding INT32_MAX so that the
 * saturated request keeps the Go meaning "more than any offer has". */
int32_t rpk_oracle_clamp_i32(int64_t v) {
    if (v > INT32_MAX) return INT32_MAX;
    if (v < INT32_MIN) return INT32_MIN;
    return (int32_t)v;
}

/* strings.ToUpper restricted to what can land on "SECURE"/"COMMUNITY":
 * ASCII a-z, plus the two non-ASCII runes whose Unicode upper case is an
 * ASCII letter (U+017F LATIN SMALL LETTER LONG S -> 'S', U+0131 DOTLESS I ->
 * 'I').  Any other non-ASCII byte can never compare equal and is copied. */
static size_t go_to_upper_ascii_target(const char *s, char *dst, size_t cap) {
    size_t o = 0;
    for (size_t i = 0; s[i] != '\0' && o + 1 < cap;) {
        unsigned char c = (unsigned char)s[i];
        if (c == 0xC5 && (unsigned char)s[i + 1] == 0xBF) { dst[o++] = 'S'; i += 2; continue; }
        if (c == 0xC4 && (unsigned char)s[i + 1] == 0xB1) { dst[o++] = 'I'; i += 2; continue; }
        if (c >= 'a' && c <= 'z') c = (unsigned char)(c - 'a' + 'A');
        dst[o++] = (char)c;
        ++i;
    }
    dst[o] = '\0';
    return o;
}

/* validateCloudType -- runpod_client.go:1115-1134: ""->SECURE; upper-case;
 * SECURE/COMMUNITY accepted; anything else (e.g. "STANDARD",
 * runpod_test.go:89) -> SECURE. */
int rpk_oracle_validate_cloud_type(const char *val) {
    if (val == NULL || val[0] == '\0') return RPK_CLOUD_SECURE;
    char up[64];
    if (strlen(val) >= sizeof(up)) return RPK_CLOUD_SECURE; /* longer than either literal even after folding */
    go_to_upper_ascii_target(val, up, sizeof(up));
    if (strcmp(up, "SECURE") == 0) return RPK_CLOUD_SECURE;
    if (strcmp(up, "COMMUNITY") == 0) return RPK_CLOUD_COMMUNITY;
    return RPK_CLOUD_SECURE;
}

/* getAnnotationWithFallback -- runpod_client.go:1102-1112: pod annotation if
 * present and non-empty, else the owner Job's, else the default.  NULL means
 * "key absent". */
const char *rpk_oracle_annotation_with_fallback(const char *pod_val, const char *job_val, const char *default_val) {
    if (pod_val != NULL && pod_val[0] != '\0') return pod_val;
    if (job_val != NULL && job_val[0] != '\0') return job_val;
    return default_val;
}

/* ------------------------------------------------------------------------ */
/* selection: GetGPUTypes -- runpod_client.go:431-520                        */
/* ------------------------------------------------------------------------ */

typedef struct {
    int32_t idx;   /* stands for ID / DisplayName (strings stay on the host, index = offer id) */
    int32_t mem;   /* MemoryInGb */
    double price;  /* Price */
} filtered_gpu;

/* sort.Slice(filteredGPUs, price asc) -- runpod_client.go:497-500.  Go's
 * sort.Slice is pdqsort: insertion sort for n<=12 (stable), unspecified tie
 * order beyond.  This repo's contract: ties -> lowest original
 * offer index, i.e. a STABLE sort, which is what Go does for n<=12 and what
 * any Go version yields on tie-free tables. */
static void insertion_sort(filtered_gpu *a, size_t n) {
    for (size_t i = 1; i < n; ++i) {
        filtered_gpu x = a[i];
        size_t j = i;
        while (j > 0 && x.price < a[j - 1].price) { a[j] = a[j - 1]; --j; }
        a[j] = x;
    }
}

static void merge_sort(filtered_gpu *a, filtered_gpu *tmp, size_t n) {
    if (n <= 12) { insertion_sort(a, n); return; }
    size_t h = n / 2;
    merge_sort(a, tmp, h);
    merge_sort(a + h, tmp, n - h);
    size_t i = 0, j = h, k = 0;
    while (i < h && j < n) tmp[k++] = (a[j].price < a[i].price) ? a[j++] : a[i++]; /* right only when strictly less: stable */
    while (i < h) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, n * sizeof(*a));
}

typedef struct {
    uint32_t G;
    const int32_t *mem_gb;          /* GPUType.MemoryInGb      runpod_client.go:86 */
    const int32_t *vcpu;            /* extension column, NULL = all 0 */
    const int32_t *ram_gb;          /* extension column, NULL = all 0 */
    const double *secure_price;     /* GPUType.SecurePrice     :88 */
    const double *community_price;  /* GPUType.CommunityPrice  :90 */
    const uint8_t *flags;           /* bit0 SecureCloud :87, bit1 CommunityCloud :89 */
} offer_table;

/* One GetGPUTypes call.  scratch/tmp hold >= G entries.  Writes up to 5
 * offer indices to out5 (-1 padded) and returns how many (0 = "No eligible
 * GPU types found", which is not an error: runpod_client.go:511-517). */
static int get_gpu_types(const offer_table *t, int64_t min_ram_per_gpu, int32_t req_vcpu, int32_t req_ram,
                         double max_price, int cloud_type, filtered_gpu *scratch, filtered_gpu *tmp,
                         int32_t *out5) {
    size_t nf = 0;
    for (uint32_t g = 0; g < t->G; ++g) { /* :465 */
        double price = 0.0;               /* :466 */
        int cloud_check = 0;              /* :467 */
        if (cloud_type == RPK_CLOUD_SECURE) { /* :469-471 */
            price = t->secure_price[g];
            cloud_check = (t->flags[g] & 1) != 0;
        } else if (cloud_type == RPK_CLOUD_COMMUNITY) { /* :472-475 */
            price = t->community_price[g];
            cloud_check = (t->flags[g] & 2) != 0;
        }
        /* :478 -- strict on price (both sides), non-strict on memory.  The
         * two extension columns follow the memory rule; with the default
         * request 0 against a table without those columns they always pass. */
        int32_t ov = t->vcpu ? t->vcpu[g] : 0;
        int32_t orm = t->ram_gb ? t->ram_gb[g] : 0;
        if (cloud_check && price > 0 && price < max_price && (int64_t)t->mem_gb[g] >= min_ram_per_gpu &&
            ov >= req_vcpu && orm >= req_ram) {
            scratch[nf].idx = (int32_t)g; /* :479-489 append */
            scratch[nf].mem = t->mem_gb[g];
            scratch[nf].price = price;
            ++nf;
        }
    }
    merge_sort(scratch, tmp, nf); /* :497-500 */
    int k = 0;
    for (size_t i = 0; i < nf; ++i) { /* :503-509 */
        if (i >= 5) break;
        out5[k++] = scratch[i].idx;
    }
    for (int i = k; i < 5; ++i) out5[i] = -1;
    return k;
}

int rpk_oracle_get_gpu_types(uint32_t G, const int32_t *mem_gb, const int32_t *vcpu, const int32_t *ram_gb,
                             const double *secure_price, const double *community_price, const uint8_t *flags,
                             int64_t min_ram_per_gpu, int32_t req_vcpu, int32_t req_ram, double max_price,
                             int cloud_type, int32_t *out5) {
    offer_table t = {G, mem_gb, vcpu, ram_gb, secure_price, community_price, flags};
    filtered_gpu *scratch = (filtered_gpu *)malloc((size_t)(G ? G : 1) * 2 * sizeof(filtered_gpu));
    if (!scratch) return -1;
    int k = get_gpu_types(&t, min_ram_per_gpu, req_vcpu, req_ram, max_price, cloud_type, scratch, scratch + (G ? G : 1), out5);
    free(scratch);
    return k;
}

/* The P x G grid as the reference would walk it: one GetGPUTypes call per
 * pod (kubelet.go:463 -> runpod_client.go:1281), pods [p0, p1). */
typedef struct {
    offer_table t;
    uint32_t p0, p1;
    const int32_t *req_mem_gb;
    const int32_t *req_vcpu;  /* NULL = 0 */
    const int32_t *req_ram_gb;/* NULL = 0 */
    const double *max_price;  /* NULL = DefaultMaxPrice 0.5, runpod_client.go:48 */
    const uint8_t *cloud;     /* NULL = SECURE */
    int32_t *best;            /* [P]   argmin = element 0 of gpuTypeIds, -1 none */
    int32_t *top5;            /* [P*5] or NULL */
    int rc;
} select_job;

static void *select_range(void *arg) {
    select_job *j = (select_job *)arg;
    uint32_t G = j->t.G;
    filtered_gpu *scratch = (filtered_gpu *)malloc((size_t)(G ? G : 1) * 2 * sizeof(filtered_gpu));
    if (!scratch) { j->rc = -1; return NULL; }
    for (uint32_t p = j->p0; p < j->p1; ++p) {
        int32_t out5[5];
        get_gpu_types(&j->t, j->req_mem_gb[p], j->req_vcpu ? j->req_vcpu[p] : 0, j->req_ram_gb ? j->req_ram_gb[p] : 0,
                      j->max_price ? j->max_price[p] : 0.5, j->cloud ? j->cloud[p] : RPK_CLOUD_SECURE, scratch,
                      scratch + (G ? G : 1), out5);
        j->best[p] = out5[0];
        if (j->top5) memcpy(j->top5 + (size_t)p * 5, out5, sizeof(out5));
    }
    free(scratch);
    j->rc = 0;
    return NULL;
}

/* n_threads <= 1: the single-goroutine shape of the reference (1 pod-sync
 * worker, cmd/virtual_kubelet/main.go:263).  n_threads > 1: one contiguous
 * pod-row range per thread, used for the all-cores CPU arm. */
int rpk_oracle_select(uint32_t G, const int32_t *mem_gb, const int32_t *vcpu, const int32_t *ram_gb,
                      const double *secure_price, const double *community_price, const uint8_t *flags, uint32_t P,
                      const int32_t *req_mem_gb, const int32_t *req_vcpu, const int32_t *req_ram_gb,
                      const double *max_price, const uint8_t *cloud, int32_t *best, int32_t *top5, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if ((uint32_t)n_threads > P && P > 0) n_threads = (int)P;
    select_job *jobs = (select_job *)calloc((size_t)n_threads, sizeof(select_job));
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    if (!jobs || !th) { free(jobs); free(th); return -1; }
    for (int i = 0; i < n_threads; ++i) {
        select_job *j = &jobs[i];
        j->t = (offer_table){G, mem_gb, vcpu, ram_gb, secure_price, community_price, flags};
        j->p0 = (uint32_t)((uint64_t)P * (uint64_t)i / (uint64_t)n_threads);
        j->p1 = (uint32_t)((uint64_t)P * (uint64_t)(i + 1) / (uint64_t)n_threads);
        j->req_mem_gb = req_mem_gb; j->req_vcpu = req_vcpu; j->req_ram_gb = req_ram_gb;
        j->max_price = max_price; j->cloud = cloud; j->best = best; j->top5 = top5;
    }
    int rc = 0;
    if (n_threads == 1) {
        select_range(&jobs[0]);
        rc = jobs[0].rc;
    } else {
        for (int i = 0; i < n_threads; ++i) pthread_create(&th[i], NULL, select_range, &jobs[i]);
        for (int i = 0; i < n_threads; ++i) { pthread_join(th[i], NULL); if (jobs[i].rc) rc = jobs[i].rc; }
    }
    free(jobs); free(th);
    return rc;
}

/* ------------------------------------------------------------------------ */
/* status sweep diff: updateAllPodStatuses -- kubelet.go:857-880             */
/* ------------------------------------------------------------------------ */
/* Canonical record: fixed `stride`-byte slot
 *   [len:u8][status ASCII ... ][0x00][ports_exposed:u8][zero pad]
 * where len counts status bytes + 2.  This holds exactly the two fields the
 * reference compares: InstanceInfo.Status and .PortsExposed
 * (runpod_client.go:103,108). */

/* statusChanged || portsExposureChanged -- kubelet.go:870-873, on decoded
 * fields (string compare + bool compare), NOT on hashes. */
static int record_changed(const uint8_t *now, const uint8_t *prev, uint32_t stride) {
    unsigned ln = now[0] & 0x7Fu, lp = prev[0] & 0x7Fu; /* bit 7 of byte 0 is the message flag: not a compared field */
    if (ln > stride - 1) ln = stride - 1; /* malformed length: clamp like the hash column does */
    if (lp > stride - 1) lp = stride - 1;
    /* decode: status = bytes[1 .. len-2], ports = bytes[len] */
    unsigned sn = ln >= 2 ? ln - 2 : 0, sp = lp >= 2 ? lp - 2 : 0;
    int status_changed = (sn != sp) || memcmp(now + 1, prev + 1, sn) != 0;          /* :870 string(status) != podInfo.Status */
    int ports_now = ln >= 2 ? (now[ln] != 0) : 0, ports_prev = lp >= 2 ? (prev[lp] != 0) : 0;
    int ports_changed = ports_now != ports_prev;                                    /* :871 */
    return status_changed || ports_changed;                                         /* :873 */
}

/* One sweep over N tracked slots.  prev is updated in place for changed rows
 * only (kubelet.go:875-880).  changed_idx is ascending.  has_prev[i]==0 marks
 * a slot that has never been seen (after a reset): it always reports changed. */
uint32_t rpk_oracle_status_diff(uint32_t N, uint32_t stride, const uint8_t *records, uint8_t *prev,
                                uint8_t *has_prev, uint32_t *changed_idx) {
    uint32_t n = 0;
    for (uint32_t i = 0; i < N; ++i) { /* :825 */
        const uint8_t *r = records + (size_t)i * stride;
        uint8_t *q = prev + (size_t)i * stride;
        if (!has_prev[i] || record_changed(r, q, stride)) {
            memcpy(q, r, stride); /* :875-880 */
            has_prev[i] = 1;
            changed_idx[n++] = i;
        }
    }
    return n;
}

/* ------------------------------------------------------------------------ */
/* XXH64 -- public xxHash specification (what cespare/xxhash/v2 Sum64         */
/* implements for seed 0; go.mod:60).  Restated from the spec.               */
/* ------------------------------------------------------------------------ */
#define P1 0x9E3779B185EBCA87ULL
#define P2 0xC2B2AE3D27D4EB4FULL
#define P3 0x165667B19E3779F9ULL
#define P4 0x85EBCA77C2B2AE63ULL
#define P5 0x27D4EB2F165667C5ULL
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; } /* little-endian host */
static inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t xround(uint64_t acc, uint64_t in) { return rotl64(acc + in * P2, 31) * P1; }
static inline uint64_t xmerge(uint64_t h, uint64_t v) { return (h ^ xround(0, v)) * P1 + P4; }

uint64_t rpk_oracle_xxh64(const uint8_t *p, size_t len, uint64_t seed) {
    const uint8_t *end = p + len;
    uint64_t h;
    if (len >= 32) {
        uint64_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
        const uint8_t *lim = end - 32;
        do {
            v1 = xround(v1, rd64(p)); v2 = xround(v2, rd64(p + 8));
            v3 = xround(v3, rd64(p + 16)); v4 = xround(v4, rd64(p + 24));
            p += 32;
        } while (p <= lim);
        h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
        h = xmerge(h, v1); h = xmerge(h, v2); h = xmerge(h, v3); h = xmerge(h, v4);
    } else {
        h = seed + P5;
    }
    h += (uint64_t)len;
    while (p + 8 <= end) { h ^= xround(0, rd64(p)); h = rotl64(h, 27) * P1 + P4; p += 8; }
    if (p + 4 <= end) { h ^= (uint64_t)rd32(p) * P1; h = rotl64(h, 23) * P2 + P3; p += 4; }
    while (p < end) { h ^= (uint64_t)(*p) * P5; h = rotl64(h, 11) * P1; ++p; }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}

/* hash column of a record table: XXH64(seed 0) over the slot's zero-padded prefix in whole 8-byte lanes --
 * bytes [0, 8*ceil((1+len)/8)) with the flag bit of byte 0 cleared (include/rpk.h) */
void rpk_oracle_record_hashes(uint32_t N, uint32_t stride, const uint8_t *records, uint64_t *out) {
    uint8_t buf[256];
    for (uint32_t i = 0; i < N; ++i) {
        const uint8_t *r = records + (size_t)i * stride;
        unsigned len = r[0] & 0x7Fu;
        if (len > stride - 1) len = stride - 1;
        unsigned nbytes = ((len + 8u) >> 3) << 3;
        if (nbytes > stride) nbytes = stride;
        memcpy(buf, r, nbytes);
        buf[0] &= 0x7Fu;
        out[i] = rpk_oracle_xxh64(buf, nbytes, 0);
    }
}

/* ------------------------------------------------------------------------ */
/* translateRunPodStatus -- kubelet.go:1848-2024, the part that depends on   */
/* (runpodStatus, statusMessage, hasExposedPorts): phase, container state,   */
/* readiness, exit code, reason, which message.  Returned as the code of     */
/* include/rpk.h (RPK_CODE_*).                                               */
/* ------------------------------------------------------------------------ */
enum { PH_UNKNOWN = 0, PH_PENDING = 1, PH_RUNNING = 2, PH_SUCCEEDED = 3, PH_FAILED = 4 };
enum { ST_WAITING = 0, ST_RUNNING = 1, ST_TERMINATED = 2 };
enum { RS_NONE = 0, RS_CONTAINER_CREATING = 1, RS_COMPLETED = 2, RS_ERROR = 3, RS_TERMINATED = 4, RS_POD_DELETED = 5, RS_STATUS_UNKNOWN = 6 };
enum { MSG_STATUS_MESSAGE = 0, MSG_PORTS_NOT_EXPOSED = 1, MSG_POD_DELETED = 2, MSG_UNKNOWN_STATUS = 3 };

static int contains_lower(const char *hay, const char *needle) { /* strings.Contains(strings.ToLower(hay), needle) */
    size_t n = strlen(needle);
    for (const char *p = hay; *p; ++p) {
        size_t k = 0;
        while (k < n && p[k] && (char)tolower((unsigned char)p[k]) == needle[k]) ++k;
        if (k == n) return 1;
    }
    return 0;
}

uint32_t rpk_oracle_translate(const char *status, const char *message, int has_exposed_ports) {
    int phase = PH_UNKNOWN, state = ST_WAITING, ready = 0, started = 0, exit_code = 0, reason = RS_NONE, msg = MSG_STATUS_MESSAGE; /* :1861-1864 */
    if (strcmp(status, "RUNNING") == 0) {                      /* :1867 */
        if (has_exposed_ports) { phase = PH_RUNNING; state = ST_RUNNING; ready = 1; started = 1; }          /* :1868-1879 */
        else { phase = PH_PENDING; state = ST_WAITING; reason = RS_CONTAINER_CREATING; msg = MSG_PORTS_NOT_EXPOSED; } /* :1880-1891 */
    } else if (strcmp(status, "STARTING") == 0) {              /* :1893 */
        phase = PH_PENDING; state = ST_WAITING; reason = RS_CONTAINER_CREATING;
    } else if (strcmp(status, "EXITED") == 0) {                /* :1905 */
        if (contains_lower(message, "error") || contains_lower(message, "fail")) { exit_code = 1; reason = RS_ERROR; phase = PH_FAILED; } /* :1909-1913 */
        else { reason = RS_COMPLETED; phase = PH_SUCCEEDED; }                                                  /* :1914-1916 */
        state = ST_TERMINATED;
    } else if (strcmp(status, "TERMINATING") == 0) {           /* :1931 */
        phase = PH_RUNNING; state = ST_RUNNING; ready = 1; started = 1;
    } else if (strcmp(status, "TERMINATED") == 0) {            /* :1943 */
        phase = PH_SUCCEEDED; state = ST_TERMINATED; reason = RS_TERMINATED;
    } else if (strcmp(status, "NOT_FOUND") == 0) {             /* :1957 */
        phase = PH_FAILED; state = ST_TERMINATED; exit_code = 1; reason = RS_POD_DELETED; msg = MSG_POD_DELETED;
    } else {                                                   /* :1971 */
        state = ST_WAITING; reason = RS_STATUS_UNKNOWN; msg = MSG_UNKNOWN_STATUS;
    }
    int ready_condition = phase == PH_RUNNING;                 /* :1982-1985; equals containerStatus.Ready in every branch */
    (void)ready;
    return (uint32_t)phase | (uint32_t)ready_condition << 3 | (uint32_t)started << 4 | (uint32_t)state << 5 | (uint32_t)exit_code << 7 |
           (uint32_t)reason << 8 | (uint32_t)msg << 11;
}

/* codes of a record table: decode (status, ports, flag) from each slot, then translate.  A set flag stands for a
 * statusMessage that contains "error"/"fail"; the string passed here exercises the case fold of :1907. */
void rpk_oracle_record_codes(uint32_t N, uint32_t stride, const uint8_t *records, uint16_t *out) {
    char status[260];
    for (uint32_t i = 0; i < N; ++i) {
        const uint8_t *r = records + (size_t)i * stride;
        unsigned len = r[0] & 0x7Fu;
        if (len > stride - 1) len = stride - 1;
        unsigned sl = len >= 2 ? len - 2 : 0;
        memcpy(status, r + 1, sl);
        status[sl] = 0;
        int ports = len >= 2 ? r[len] != 0 : 0;
        out[i] = (uint16_t)rpk_oracle_translate(status, (r[0] & 0x80u) ? "Container FAILED: exit status 1" : "", ports);
    }
}

/* Reference-shaped sweep over Go-like rows, used only as the timed CPU arm:
 * status strings compared as strings, ports as bools (kubelet.go:870-871). */
uint32_t rpk_oracle_status_sweep_strings(uint32_t N, const char *const *status_now, const uint8_t *ports_now,
                                         const char **status_prev, uint8_t *ports_prev, uint32_t *changed_idx) {
    uint32_t n = 0;
    for (uint32_t i = 0; i < N; ++i) {
        int sc = strcmp(status_now[i], status_prev[i]) != 0;
        int pc = (ports_now[i] != 0) != (ports_prev[i] != 0);
        if (sc || pc) { status_prev[i] = status_now[i]; ports_prev[i] = ports_now[i]; changed_idx[n++] = i; }
    }
    return n;
}
