/* oracle/nms_ref.c — TEST INFRASTRUCTURE, not product code.
 *
 * Scalar CPU restatement of the arithmetic the reference delegates to torchvision.ops.nms
 * (call site: yololite/utils/ops.py:265; torchvision 0.26.0 CPU kernel `nms_kernel_impl`, which is NOT in
 * /root/reference — un-vendored third-party dependency, version unpinned by the reference).  Published
 * algorithm restated here: stable descending sort of the scores, areas (x2-x1)*(y2-y1) in fp32, greedy scan
 * suppressing j when  inter / (area_i + area_j - inter)  (fp32, left-to-right, no FMA)  >  iou_threshold
 * (compared in double).  Pinned by tests/golden/nms_torchvision_*.npz = outputs of the installed torchvision
 * CPU kernel (oracle/gen_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: contraction would change the low bits).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    float score;
    int32_t idx;
} item_t;

/* stable descending merge sort (qsort is not stable) */
static void merge_sort(item_t* a, item_t* tmp, int n) {
    if (n < 2) return;
    int h = n / 2;
    merge_sort(a, tmp, h);
    merge_sort(a + h, tmp, n - h);
    int i = 0, j = h, k = 0;
    while (i < h && j < n) {
        /* take from the right run only when strictly greater: equal scores keep original order */
        if (a[j].score > a[i].score) tmp[k++] = a[j++];
        else tmp[k++] = a[i++];
    }
    while (i < h) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, (size_t)n * sizeof(item_t));
}

/* boxes: n x 4 (x1,y1,x2,y2) fp32; keep: out, int64[n]; returns number kept.  max_keep <= 0: no limit. */
int64_t yl_ref_nms(const float* boxes, const float* scores, int64_t n, double iou_threshold, int64_t* keep,
                   int64_t max_keep) {
    if (n <= 0) return 0;
    item_t* order = (item_t*)malloc((size_t)n * sizeof(item_t));
    item_t* tmp = (item_t*)malloc((size_t)n * sizeof(item_t));
    float* areas = (float*)malloc((size_t)n * sizeof(float));
    uint8_t* suppressed = (uint8_t*)calloc((size_t)n, 1);
    for (int64_t i = 0; i < n; ++i) {
        order[i].score = scores[i];
        order[i].idx = (int32_t)i;
        areas[i] = (boxes[4 * i + 2] - boxes[4 * i + 0]) * (boxes[4 * i + 3] - boxes[4 * i + 1]);
    }
    merge_sort(order, tmp, (int)n);
    int64_t num = 0;
    for (int64_t _i = 0; _i < n; ++_i) {
        int64_t i = order[_i].idx;
        if (suppressed[i]) continue;
        keep[num++] = i;
        if (max_keep > 0 && num >= max_keep) break; /* later boxes cannot change earlier decisions */
        float ix1 = boxes[4 * i + 0], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3];
        float iarea = areas[i];
        for (int64_t _j = _i + 1; _j < n; ++_j) {
            int64_t j = order[_j].idx;
            if (suppressed[j]) continue;
            float xx1 = ix1 < boxes[4 * j + 0] ? boxes[4 * j + 0] : ix1; /* std::max(ix1, x1[j]) */
            float yy1 = iy1 < boxes[4 * j + 1] ? boxes[4 * j + 1] : iy1;
            float xx2 = boxes[4 * j + 2] < ix2 ? boxes[4 * j + 2] : ix2; /* std::min(ix2, x2[j]) */
            float yy2 = boxes[4 * j + 3] < iy2 ? boxes[4 * j + 3] : iy2;
            float dw = xx2 - xx1, dh = yy2 - yy1;
            float w = 0.0f < dw ? dw : 0.0f; /* std::max(0, xx2 - xx1) */
            float h = 0.0f < dh ? dh : 0.0f;
            float inter = w * h;
            float ovr = inter / (iarea + areas[j] - inter);
            if ((double)ovr > iou_threshold) suppressed[j] = 1;
        }
    }
    free(order);
    free(tmp);
    free(areas);
    free(suppressed);
    return num;
}
