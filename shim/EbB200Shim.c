/*
 * shim/EbB200Shim.c -- see EbB200Shim.h.  Our own code (not reference code); it includes the reference's headers because it lives inside the
 * reference's library once shim/build_shim.sh has put it there.
 */
#include <dlfcn.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "EbB200Shim.h"
#include "EbReferenceObject.h"
#include "EbMotionEstimationProcess.h"
#include "svt_hevc_b200.h"

typedef int (*fn_init)(int);
typedef int (*fn_me_picture)(const svtb200_me_config *, const svtb200_pa_picture *, const svtb200_pa_picture *, const svtb200_pa_picture *, svtb200_me_pu_result *,
                             uint32_t *, uint32_t *, uint32_t *);
typedef int (*fn_ois_picture)(const svtb200_ois_config *, const svtb200_plane *, const svtb200_me_pu_result *, uint8_t *, uint32_t *, uint32_t);
typedef const char *(*fn_last_error)(void);

static struct {
    int mode;                 /* -1 not read yet, 0 off, 1 on */
    void *lib;
    fn_init init;
    fn_me_picture me_picture;
    fn_ois_picture ois_picture;
    fn_last_error last_error;
    int ois;                  /* SVT_HEVC_B200_OIS=0 keeps the open-loop intra search on the CPU */
    unsigned long long me_pictures, me_lcus, ois_pictures;
} g = { -1, NULL, NULL, NULL, NULL, NULL, 1, 0, 0, 0 };

int EbB200Init(void)
{
    const char *e = getenv("SVT_HEVC_B200");
    g.mode = (e && e[0] && strcmp(e, "0") != 0) ? 1 : 0;
    if (!g.mode) return 0;
    if (g.lib) return 0;
    const char *path = getenv("SVT_HEVC_B200_LIB");
    if (!path || !path[0]) path = "/root/repo/svt-hevc_b200/libsvthevc_b200.so";
    g.lib = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!g.lib) { fprintf(stderr, "SVT [B200]: cannot load %s: %s\n", path, dlerror()); return -1; }
    g.init = (fn_init)dlsym(g.lib, "svtb200_init");
    g.me_picture = (fn_me_picture)dlsym(g.lib, "svtb200_me_picture");
    g.ois_picture = (fn_ois_picture)dlsym(g.lib, "svtb200_ois_picture");
    g.last_error = (fn_last_error)dlsym(g.lib, "svtb200_last_error");
    { const char *o = getenv("SVT_HEVC_B200_OIS"); g.ois = !(o && strcmp(o, "0") == 0); }
    if (!g.init || !g.me_picture || !g.ois_picture || !g.last_error) { fprintf(stderr, "SVT [B200]: %s lacks the expected entry points\n", path); return -1; }
    const char *dev = getenv("SVT_HEVC_B200_DEVICE");
    const int rc = g.init(dev ? atoi(dev) : 0);
    if (rc) { fprintf(stderr, "SVT [B200]: svtb200_init failed (%d): %s\n", rc, g.last_error()); return -1; }
    fprintf(stderr, "SVT [B200]: motion estimation%s on the GPU (%s)\n", g.ois ? " and the open-loop intra search run" : " runs", path);
    return 0;
}

int EbB200MeActive(const SequenceControlSet_t *scs, const PictureParentControlSet_t *pcs, const MeContext_t *me)
{
    if (g.mode != 1) return 0;
    if (!scs->staticConfig.unrestrictedMotionVector || scs->staticConfig.tileRowCount * scs->staticConfig.tileColumnCount != 1) return 0;   /* tile-clamped vectors: CPU */
    if (me->fractionalSearchMethod == SSD_SEARCH || me->fractionalSearch64x64) return 0;   /* enc modes <= 4 / <= 2: not built in the library */
    if (pcs->hierarchicalLevels > 5) return 0;
    return 1;
}

static void plane_of(svtb200_plane *p, const EbPictureBufferDesc_t *d)
{
    p->buf = d->bufferY; p->stride = d->strideY; p->origin_x = d->originX; p->origin_y = d->originY; p->width = d->width; p->height = d->height;
}
static void pa_of(svtb200_pa_picture *o, const EbPictureBufferDesc_t *full, const EbPictureBufferDesc_t *quarter, const EbPictureBufferDesc_t *sixteenth)
{
    memset(o, 0, sizeof *o);
    plane_of(&o->full, full);
    if (quarter) plane_of(&o->quarter, quarter);
    if (sixteenth) plane_of(&o->sixteenth, sixteenth);
}

/* returns the library's result records (caller frees): the open-loop intra search of the same picture reads their distortions */
static svtb200_me_pu_result *me_picture(SequenceControlSet_t *scs, PictureParentControlSet_t *pcs, MeContext_t *me, EbPictureBufferDesc_t *inputPadded,
                                        EbPictureBufferDesc_t *quarter, EbPictureBufferDesc_t *sixteenth)
{
    svtb200_me_config c;
    memset(&c, 0, sizeof c);
    c.pic_width = scs->lumaWidth; c.pic_height = scs->lumaHeight;
    c.num_lists = pcs->sliceType == EB_P_PICTURE ? 1 : 2;
    c.temporal_layer = pcs->temporalLayerIndex; c.hierarchical_levels = pcs->hierarchicalLevels;
    c.enable_hme = pcs->enableHmeFlag; c.enable_hme_l0 = pcs->enableHmeLevel0Flag; c.enable_hme_l1 = pcs->enableHmeLevel1Flag; c.enable_hme_l2 = pcs->enableHmeLevel2Flag;
    c.one_quadrant_hme = me->oneQuadrantHME; c.update_hme_search_center = me->updateHmeSearchCenter;
    c.n_hme_regions_w = (uint8_t)me->numberHmeSearchRegionInWidth; c.n_hme_regions_h = (uint8_t)me->numberHmeSearchRegionInHeight;
    c.hme_l0_total_w = me->hmeLevel0TotalSearchAreaWidth; c.hme_l0_total_h = me->hmeLevel0TotalSearchAreaHeight;
    for (int i = 0; i < 2; ++i) {
        c.hme_l0_w[i] = me->hmeLevel0SearchAreaInWidthArray[i]; c.hme_l0_h[i] = me->hmeLevel0SearchAreaInHeightArray[i];
        c.hme_l1_w[i] = me->hmeLevel1SearchAreaInWidthArray[i]; c.hme_l1_h[i] = me->hmeLevel1SearchAreaInHeightArray[i];
        c.hme_l2_w[i] = me->hmeLevel2SearchAreaInWidthArray[i]; c.hme_l2_h[i] = me->hmeLevel2SearchAreaInHeightArray[i];
        c.ref_poc[i] = pcs->refPicPocArray[i];
    }
    c.search_area_w = me->searchAreaWidth; c.search_area_h = me->searchAreaHeight;
    c.fractional_search_model = me->fractionalSearchModel; c.fractional_search_method = me->fractionalSearchMethod;
    c.cu8x8_mode = pcs->cu8x8Mode; c.cu16x16_mode = pcs->cu16x16Mode;
    c.rate_control_mode = (uint8_t)scs->staticConfig.rateControlMode;
    c.lambda = me->lambda;
    for (int i = 0; i < 12; ++i) c.mvd_bits[i] = me->mvdBitsArray[i];

    svtb200_pa_picture src, ref[2];
    pa_of(&src, inputPadded, quarter, sixteenth);
    for (int l = 0; l < c.num_lists; ++l) {
        EbPaReferenceObject_t *r = (EbPaReferenceObject_t *)pcs->refPaPicPtrArray[l]->objectPtr;
        pa_of(&ref[l], r->inputPaddedPicturePtr, r->quarterDecimatedPicturePtr, r->sixteenthDecimatedPicturePtr);
    }
    const EB_U32 cols = (scs->lumaWidth + 63) / 64, rows = (scs->lumaHeight + 63) / 64, n = cols * rows;
    svtb200_me_pu_result *res = (svtb200_me_pu_result *)malloc((size_t)n * SVTB200_ME_PU_COUNT * sizeof(*res));
    uint32_t *rcd = (uint32_t *)malloc((size_t)n * sizeof(uint32_t));
    const int rc = (res && rcd) ? g.me_picture(&c, &src, &ref[0], c.num_lists == 2 ? &ref[1] : NULL, res, NULL, NULL, rcd) : -100;
    if (rc) {
        /* the library promises never to fail at call time once EbInitEncoder succeeded (SURVEY 8b "Errors"); if it does, stop like the reference's own
         * fatal-error path (EbErrorHandling.h:15) rather than continue with a different encoder */
        fprintf(stderr, "SVT [B200]: svtb200_me_picture failed (%d): %s\n", rc, g.last_error ? g.last_error() : "?");
        abort();
    }
    const EB_U32 npu = pcs->maxNumberOfPusPerLcu < SVTB200_ME_PU_COUNT ? pcs->maxNumberOfPusPerLcu : SVTB200_ME_PU_COUNT;
    for (EB_U32 lcu = 0; lcu < n; ++lcu) {
        for (EB_U32 p = 0; p < npu; ++p) {
            const svtb200_me_pu_result *r = &res[(size_t)lcu * SVTB200_ME_PU_COUNT + p];
            MeCuResults_t *o = &pcs->meResults[lcu][p];
            o->xMvL0 = r->x_mv_l0; o->yMvL0 = r->y_mv_l0; o->xMvL1 = r->x_mv_l1; o->yMvL1 = r->y_mv_l1;
            for (int k = 0; k < 3; ++k) { o->distortionDirection[k].distortion = r->distortion[k]; o->distortionDirection[k].direction = r->direction[k]; }
            o->totalMeCandidateIndex = r->total_me_candidates;
        }
        pcs->rcMEdistortion[lcu] = rcd[lcu];
    }
    free(rcd);
    __sync_fetch_and_add(&g.me_pictures, 1ULL);
    __sync_fetch_and_add(&g.me_lcus, (unsigned long long)n);
    return res;
}

/* OpenLoopIntraSearchLcu of every LCU (Codec/EbMotionEstimation.c:5053): the library returns, per CU, the number of candidates and the candidates as the
 * reference's own packed OisCandidate_t word (distortion | validDistortion << 20 | intraMode << 24) */
static void ois_picture(SequenceControlSet_t *scs, PictureParentControlSet_t *pcs, MotionEstimationContext_t *ctx, EbPictureBufferDesc_t *inputPicture,
                        const svtb200_me_pu_result *me_res)
{
    svtb200_ois_config c;
    memset(&c, 0, sizeof c);
    c.pic_width = scs->lumaWidth; c.pic_height = scs->lumaHeight;
    c.slice_type = pcs->sliceType == EB_I_PICTURE ? 2 : pcs->sliceType == EB_P_PICTURE ? 1 : 0;
    c.temporal_layer = pcs->temporalLayerIndex;
    c.ois_kernel_level = ctx->oisKernelLevel; c.ois_th_set = ctx->oisThSet; c.set_best_ois_distortion_to_valid = ctx->setBestOisDistortionToValid;
    c.limit_ois_to_dc_mode = pcs->limitOisToDcModeFlag; c.skip_ois_8x8 = pcs->skipOis8x8; c.cu8x8_mode = pcs->cu8x8Mode;
    svtb200_plane src;
    plane_of(&src, inputPicture);
    const EB_U32 cols = (scs->lumaWidth + 63) / 64, rows = (scs->lumaHeight + 63) / 64, n = cols * rows;
    const uint32_t cap = c.ois_kernel_level ? 18 : 9;
    uint8_t *total = (uint8_t *)malloc((size_t)n * 85);
    uint32_t *cand = (uint32_t *)malloc((size_t)n * 85 * cap * sizeof(uint32_t));
    const int rc = (total && cand) ? g.ois_picture(&c, &src, me_res, total, cand, cap) : -100;
    if (rc) { fprintf(stderr, "SVT [B200]: svtb200_ois_picture failed (%d): %s\n", rc, g.last_error ? g.last_error() : "?"); abort(); }
    for (EB_U32 lcu = 0; lcu < n; ++lcu)
        for (int cu = 1; cu < 85; ++cu) {
            const uint8_t t = total[(size_t)lcu * 85 + cu];
            OisCandidate_t *o;
            if (cu < 21) { pcs->oisCu32Cu16Results[lcu]->totalIntraLumaMode[cu] = t; o = pcs->oisCu32Cu16Results[lcu]->sortedOisCandidate[cu]; }
            else { pcs->oisCu8Results[lcu]->totalIntraLumaMode[cu - 21] = t; o = pcs->oisCu8Results[lcu]->sortedOisCandidate[cu - 21]; }
            for (uint32_t k = 0; k < cap && k < (uint32_t)(t ? t : 1); ++k) o[k].oisResults = cand[((size_t)lcu * 85 + cu) * cap + k];
        }
    free(total); free(cand);
    __sync_fetch_and_add(&g.ois_pictures, 1ULL);
}

/* one run per picture, whichever ME segment thread arrives first */
#define B200_SLOTS 256
static struct { const void *pcs; EB_U64 poc; int state, did; } g_slot[B200_SLOTS];
static unsigned g_next;
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_cv = PTHREAD_COND_INITIALIZER;

int EbB200PictureOnce(SequenceControlSet_t *scs, PictureParentControlSet_t *pcs, MotionEstimationContext_t *ctx, EbPictureBufferDesc_t *inputPicture,
                      EbPictureBufferDesc_t *inputPadded, EbPictureBufferDesc_t *quarter, EbPictureBufferDesc_t *sixteenth)
{
    if (g.mode != 1) return 0;
    int mine = -1, did;
    pthread_mutex_lock(&g_mu);
    for (int i = 0; i < B200_SLOTS; ++i)
        if (g_slot[i].state && g_slot[i].pcs == (const void *)pcs && g_slot[i].poc == pcs->pictureNumber) { mine = i; break; }
    if (mine < 0) {
        mine = (int)(g_next++ % B200_SLOTS);
        g_slot[mine].pcs = pcs; g_slot[mine].poc = pcs->pictureNumber; g_slot[mine].state = 1; g_slot[mine].did = 0;
        pthread_mutex_unlock(&g_mu);
        did = 0;
        svtb200_me_pu_result *res = NULL;
        if (pcs->sliceType != EB_I_PICTURE && EbB200MeActive(scs, pcs, ctx->meContextPtr)) {
            res = me_picture(scs, pcs, ctx->meContextPtr, inputPadded, quarter, sixteenth);
            did |= EB_B200_DID_ME;
        }
        /* the intra search reads the ME distortions of the same picture: on non-I pictures it goes to the GPU only together with ME */
        if (g.ois && (pcs->sliceType == EB_I_PICTURE || res)) {
            ois_picture(scs, pcs, ctx, inputPicture, res);
            did |= EB_B200_DID_OIS;
        }
        free(res);
        pthread_mutex_lock(&g_mu);
        g_slot[mine].did = did; g_slot[mine].state = 2;
        pthread_cond_broadcast(&g_cv);
        pthread_mutex_unlock(&g_mu);
        return did;
    }
    while (g_slot[mine].state == 1 && g_slot[mine].pcs == (const void *)pcs && g_slot[mine].poc == pcs->pictureNumber) pthread_cond_wait(&g_cv, &g_mu);
    did = g_slot[mine].did;
    pthread_mutex_unlock(&g_mu);
    return did;
}

void EbB200Report(void)
{
    if (g.mode == 1) fprintf(stderr, "SVT [B200]: %llu pictures / %llu LCUs of motion estimation and %llu pictures of open-loop intra search ran on the GPU\n", g.me_pictures, g.me_lcus, g.ois_pictures);
}
