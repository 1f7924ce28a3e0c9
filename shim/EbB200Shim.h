/*
 * shim/EbB200Shim.h -- the reference-side half of the drop-in boundary.  Compiled INTO a scratch copy of the reference's
 * Source/Lib/Codec by shim/build_shim.sh together with shim/b200.patch; nothing here changes Source/API (EbApi.h, EbInitEncoder, the app and the
 * ffmpeg / gstreamer plugins stay byte-identical).
 *
 * Backend selection is by environment, next to the ASM_TYPES detection of EbInitEncoder (Codec/EbEncHandle.c:702-707):
 *   SVT_HEVC_B200 unset / "0"   the reference's own C / SIMD path (function-pointer tables indexed by ASM_TYPES)
 *   SVT_HEVC_B200 = "1"         picture-level calls into libsvthevc_b200.so (dlopen; SVT_HEVC_B200_LIB overrides its path); a missing library
 *                               or GPU makes EbInitEncoder fail with EB_ErrorInsufficientResources -- there is no silent CPU fallback
 */
#ifndef EbB200Shim_h
#define EbB200Shim_h

#include "EbDefinitions.h"
#include "EbPictureControlSet.h"
#include "EbSequenceControlSet.h"
#include "EbMotionEstimationContext.h"
#include "EbPictureBufferDesc.h"

#ifdef __cplusplus
extern "C" {
#endif

/* EbInitEncoder hook: 0 = ok (backend off or initialised), else the caller returns EB_ErrorInsufficientResources */
extern int EbB200Init(void);
/* 1 when motion estimation of this picture goes to the GPU (backend on, unrestricted motion vectors, one tile, a preset the library supports) */
extern int EbB200MeActive(const SequenceControlSet_t *scs, const PictureParentControlSet_t *pcs, const MeContext_t *me);
/* The open-loop stages of one picture in one go -- MotionEstimateLcu of EVERY LCU (fills pcs->meResults[lcu][pu] and pcs->rcMEdistortion[lcu], what
 * MotionEstimateLcu writes, Codec/EbMotionEstimation.c:4321-4438) and OpenLoopIntraSearchLcu of every LCU (pcs->oisCu32Cu16Results / oisCu8Results,
 * :5053).  Called by every ME segment thread of the picture (Codec/EbMotionEstimationProcess.c:706); the first caller runs the picture, the others wait
 * for it.  Returns a mask of what the GPU did: EB_B200_DID_ME, EB_B200_DID_OIS (0: backend off). */
#define EB_B200_DID_ME  1
#define EB_B200_DID_OIS 2
struct MotionEstimationContext_s;
extern int EbB200PictureOnce(SequenceControlSet_t *scs, PictureParentControlSet_t *pcs, struct MotionEstimationContext_s *ctx, EbPictureBufferDesc_t *inputPicture,
                             EbPictureBufferDesc_t *inputPadded, EbPictureBufferDesc_t *quarter, EbPictureBufferDesc_t *sixteenth);
/* statistics for the log at encoder shutdown */
extern void EbB200Report(void);

#ifdef __cplusplus
}
#endif
#endif
