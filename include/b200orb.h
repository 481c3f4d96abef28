/* b200orb -- C ABI of the B200-native stereo ORB front-end.
 *
 * Drop-in boundary for ONE hot path of M2219/pyOrbSLAM: ORB extraction of the left and right image plus
 * Frame.compute_stereo_matches.  Every entry point names the reference interface it replaces
 * (paths relative to the reference root).  Plain pointers and sizes only; no torch / pybind types.
 *
 * Conventions
 *   - every function returns 0 on success, a negative B200ORB_E_* code on failure;
 *     b200orb_last_error() returns the message of the last failure on the calling thread.
 *   - "kps" arrays are float[n][6] = (x, y, size, angle, response, octave): the tuple layout of the
 *     reference's cv::KeyPoint caster (pyORBExtractor/opencv_type_casters.h:107).
 *   - "desc" arrays are uint8[n][32] (256-bit rBRIEF), row i belongs to keypoint i
 *     (pyORBExtractor/ORBextractor.cpp:1067,1087-1088).
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with B200ORB_E_CUDA.
 */
#ifndef B200ORB_H
#define B200ORB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200ORB_OK 0
#define B200ORB_E_ARG (-1)      /* bad argument / unsupported geometry */
#define B200ORB_E_CUDA (-2)     /* CUDA runtime error (incl. no device) */
#define B200ORB_E_STATE (-3)    /* call order violated (e.g. results requested before extract) */
#define B200ORB_E_RANGE (-4)    /* stereo: a keypoint row / SAD window leaves the pyramid (reference raises) */

#define B200ORB_MAX_LEVELS 16

const char* b200orb_last_error(void);
int b200orb_version(void);
/* number of CUDA devices visible; <= 0 means the product cannot run here */
int b200orb_device_count(void);
/* kernels launched by this library since load (all contexts); bench.py reports the delta as gpu_launches */
long long b200orb_kernel_launches(void);

/* ---------------------------------------------------------------------------------------------
 * Single-image extractor object == ORB_SLAM2::ORBextractor
 * (pyORBExtractor/ORBextractor.h:45-113, bound at pyORBExtractor/orb_extractor.cpp:22-38)
 * ------------------------------------------------------------------------------------------- */
typedef struct b200orb_extractor b200orb_extractor;

/* ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)  -- ORBextractor.cpp:410-470 */
int b200orb_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST,
                             int device, b200orb_extractor** out);
void b200orb_extractor_destroy(b200orb_extractor* e);

/* GetLevels / GetScaleFactor / GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares /
 * GetInverseScaleSigmaSquares  -- ORBextractor.h:62-82; each array has GetLevels() entries */
int b200orb_get_levels(const b200orb_extractor* e);
float b200orb_get_scale_factor(const b200orb_extractor* e);
int b200orb_get_scale_factors(const b200orb_extractor* e, float* out);
int b200orb_get_inverse_scale_factors(const b200orb_extractor* e, float* out);
int b200orb_get_scale_sigma_squares(const b200orb_extractor* e, float* out);
int b200orb_get_inverse_scale_sigma_squares(const b200orb_extractor* e, float* out);
/* per-level feature quota (mnFeaturesPerLevel, ORBextractor.cpp:435-446) -- diagnostic */
int b200orb_get_features_per_level(const b200orb_extractor* e, int* out);

/* operator_kd(image) -- ORBextractor.cpp:1042-1104 via orb_extractor.cpp:31-38.
 * image: host uint8[H][W], C-contiguous (CV_8UC1).  Writes the keypoint count to *n_keypoints and keeps
 * the results (and the pyramid) resident on the device until the next call.  H == 0 or W == 0 -> 0 keypoints. */
int b200orb_extract(b200orb_extractor* e, const uint8_t* image, int H, int W, int* n_keypoints);
/* copies of the last results: kps float[n][6], desc uint8[n][32] (either may be NULL) */
int b200orb_get_results(b200orb_extractor* e, float* kps, uint8_t* desc);
/* upper bound of keypoints one call can return for the current image size (>= nfeatures) */
int b200orb_max_keypoints(const b200orb_extractor* e);

/* GetImagePyramid() -- ORBextractor.h:84-86 through the Mat->ndarray caster, which copies rows*cols
 * CONTIGUOUS bytes from the level's ROI start inside its (w+38)-pitch bordered buffer
 * (opencv_type_casters.h:230-239): out[r*w + c] = bordered.flat[19*(w+38) + 19 + r*w + c].
 * That sheared view is what Frame.compute_stereo_matches reads; we return exactly it. */
int b200orb_level_size(const b200orb_extractor* e, int level, int* w, int* h);
int b200orb_get_pyramid_level(b200orb_extractor* e, int level, uint8_t* out /* h*w bytes */);
/* all levels of GetImagePyramid() in one call: the views of level 0, 1, ... concatenated (sum of h_l * w_l bytes, which must
 * be <= cap); one device synchronisation instead of one per level */
int b200orb_get_pyramid_all(b200orb_extractor* e, uint8_t* out, long long cap);
/* the true level image (dense h*w copy of the ROI) and its 7x7 sigma-2 blur -- diagnostics / parity tests */
int b200orb_get_level_image(b200orb_extractor* e, int level, int blurred, uint8_t* out /* h*w bytes */);
/* FAST candidates fed to DistributeOctTree for one level, int[cap][3] = (x, y, response), coordinates
 * relative to the (16,16) detection origin, in the reference's order (ORBextractor.cpp:788-828); returns count */
int b200orb_get_level_candidates(b200orb_extractor* e, int level, int cap, int* out, int* n);

/* ---------------------------------------------------------------------------------------------
 * Frame.compute_stereo_matches -- Frame.py:161-279
 * ------------------------------------------------------------------------------------------- */
/* device-resident form: uses the keypoints, descriptors and pyramids left on the device by the last
 * b200orb_extract of `left` and `right` (Frame.__init__ order: ExtractORB x2, then compute_stereo_matches,
 * Frame.py:48-65).  mbf = Camera.bf (python float), fx = mK[0][0] (float32).
 * uRight/depth: float[nLeft], -1 = no match (Frame.py:163-164,277-278).  matchIdx (optional): index of the
 * Hamming winner in the right image, -1 if bestDist >= (TH_HIGH+TH_LOW)/2 (Frame.py:203-222). */
int b200orb_stereo(b200orb_extractor* left, b200orb_extractor* right, double mbf, float fx,
                   float* uRight, float* depth, int* matchIdx);

/* Same with options and the SAD minimum of every accepted match (sadDist, optional, -1 = no match).
 * flags = 0 is exactly b200orb_stereo, i.e. exactly the reference.  The two flags are opt-in extensions for users who
 * want upstream ORB-SLAM2 behaviour instead; the reference does NOT behave this way:
 *   B200ORB_STEREO_MEDIAN_CULL   drop matches whose SAD minimum is >= 1.5f*1.4f*median (upstream ComputeStereoMatches;
 *                                in the reference vDistIdx is filled and then discarded, Frame.py:185,279)
 *   B200ORB_STEREO_DENSE_PYRAMID take SAD windows from the true level image instead of the step-ignoring view the
 *                                reference's Mat caster hands to Python (opencv_type_casters.h:230-239) */
#define B200ORB_STEREO_MEDIAN_CULL 1
#define B200ORB_STEREO_DENSE_PYRAMID 2
int b200orb_stereo_ex(b200orb_extractor* left, b200orb_extractor* right, double mbf, float fx, int flags,
                      float* uRight, float* depth, int* matchIdx, int* sadDist);

/* general form on caller-supplied host data (any keypoints, e.g. the stereo-only sweep):
 * kps*: float[n][3] = (x, y, octave); pyr*: nlevels pointers to the GetImagePyramid() views, level l is
 * uint8[lh[l]][lw[l]]; sf/isf: GetScaleFactors()/GetInverseScaleFactors(). */
int b200orb_stereo_host(int device, int nLeft, const float* kpsL, const uint8_t* descL,
                        int nRight, const float* kpsR, const uint8_t* descR,
                        int nlevels, const float* sf, const float* isf,
                        const uint8_t* const* pyrL, const uint8_t* const* pyrR, const int* lw, const int* lh,
                        double mbf, float fx, float* uRight, float* depth, int* matchIdx);

/* ---------------------------------------------------------------------------------------------
 * Batched stereo front-end (throughput API): what Tracking.grab_image_stereo -> Frame.__init__ does for
 * one pair (Tracking.py:95-112, Frame.py:48-65), for n_pairs independent pairs per call.
 * ------------------------------------------------------------------------------------------- */
typedef struct b200orb_batch b200orb_batch;

int b200orb_batch_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST,
                         int H, int W, int max_pairs, int device, b200orb_batch** out);
void b200orb_batch_destroy(b200orb_batch* b);
int b200orb_batch_max_pairs(const b200orb_batch* b);
/* row capacity of every per-image output array (>= any possible keypoint count) */
int b200orb_batch_kp_capacity(const b200orb_batch* b);
/* bytes of device workspace held by the batch object */
long long b200orb_batch_workspace_bytes(const b200orb_batch* b);

/* Inputs already in device memory: left/right uint8[n_pairs][H][W].  Outputs in device memory, caller-
 * allocated with C = b200orb_batch_kp_capacity():
 *   kps   float[2][n_pairs][C][6]   (index 0 = left images, 1 = right images)
 *   desc  uint8[2][n_pairs][C][32]
 *   nkp   int32[2][n_pairs]
 *   uRight, depth  float[n_pairs][C];  matchIdx int32[n_pairs][C]   (entries >= nkp[0][p] are unspecified)
 * Asynchronous on `stream` (a cudaStream_t, may be NULL for the default stream). */
int b200orb_batch_run_device(b200orb_batch* b, const uint8_t* d_left, const uint8_t* d_right, int n_pairs,
                             double mbf, float fx, float* d_kps, uint8_t* d_desc, int32_t* d_nkp,
                             float* d_uRight, float* d_depth, int32_t* d_matchIdx, void* stream);

/* Same work with HOST buffers (pinned recommended): chunks of max_pairs pairs are uploaded, processed and
 * downloaded with copy/compute overlap.  n_pairs may exceed max_pairs.  Host outputs use the device layout
 * with n_pairs = the whole job: kps float[2][n_pairs][C][6], desc uint8[2][n_pairs][C][32], nkp int32[2][n_pairs],
 * uRight/depth float[n_pairs][C], matchIdx int32[n_pairs][C] (matchIdx may be NULL).
 * Synchronous: returns when all outputs are in host memory. */
int b200orb_batch_run_host(b200orb_batch* b, const uint8_t* h_left, const uint8_t* h_right, int n_pairs,
                           double mbf, float fx, float* h_kps, uint8_t* h_desc, int32_t* h_nkp,
                           float* h_uRight, float* h_depth, int32_t* h_matchIdx);

/* One shard of a larger job (frames are independent, SURVEY.md 8e: contiguous frame ranges per GPU, no collective): processes
 * pairs [first_pair, first_pair + n_pairs) of a job of job_pairs pairs.  h_left / h_right point at THIS shard's images; the output
 * pointers are the JOB's arrays (layout above with n_pairs = job_pairs) and only this shard's rows are written, so several engines
 * -- one per GPU, each called from its own host thread -- fill one set of result arrays.  b200orb_batch_run_host is the shard
 * (job_pairs = n_pairs, first_pair = 0).  pyorbslam_b200.StereoFrontendMulti is the host side of this. */
int b200orb_batch_run_host_shard(b200orb_batch* b, const uint8_t* h_left, const uint8_t* h_right, int n_pairs,
                                 double mbf, float fx, float* h_kps, uint8_t* h_desc, int32_t* h_nkp,
                                 float* h_uRight, float* h_depth, int32_t* h_matchIdx, int job_pairs, int first_pair);

/* Range errors of the batched path.  Frame.compute_stereo_matches raises (IndexError / ValueError, Frame.py:192,230-250) when a
 * keypoint's row band or SAD window leaves the pyramid view; b200orb_stereo reports that as B200ORB_E_RANGE.  The batched calls keep one
 * flag word per pair (0 = fine): b200orb_batch_run_host returns B200ORB_E_RANGE after all outputs have reached host memory if any
 * pair of the job is flagged (the other pairs' results are valid; a flagged pair holds -1 at the offending keypoints), and
 * b200orb_batch_status_host copies the job's per-pair flags.  b200orb_batch_run_device is asynchronous, so its caller asks:
 * b200orb_batch_status_device synchronises `stream`, copies the flags of the last run_device call (pair_status may be NULL) and
 * returns B200ORB_E_RANGE if any is set. */
int b200orb_batch_status_device(b200orb_batch* b, int n_pairs, void* stream, int32_t* pair_status);
int b200orb_batch_status_host(const b200orb_batch* b, int32_t* pair_status, int n_pairs);

/* measurement aid: with on != 0, b200orb_batch_run_host / _shard perform exactly their uploads, downloads, stream waits and events but
 * launch no kernel -- the ceiling the host <-> device links of the box set for this traffic pattern (bench.py reports e2e against it) */
int b200orb_batch_set_copy_only(b200orb_batch* b, int on);

/* stereo options of the batched path (same flags as b200orb_stereo_ex; default 0 = the reference's behaviour) */
int b200orb_batch_set_stereo_flags(b200orb_batch* b, int flags);

/* diagnostic: total number of FAST candidates (the octree kernel's input) in the first n_images slots of the last
 * run -- bench.py uses it for the octree kernel's algorithmic byte count */
int b200orb_batch_candidate_count(b200orb_batch* b, int n_images, long long* total);

/* Per-kernel timing of the batched path, measured with CUDA events recorded on the launching stream between
 * the kernels of every b200orb_batch_run_device call (up to max_calls calls are kept).  Stages:
 * 0 level-0 border, 1 resize chain (nlevels-1 launches), 2 blur, 3 FAST cells, 4 octree, 5 orient+describe, 6 stereo.
 * _read sums the elapsed milliseconds per stage over the recorded calls, reports how many calls / pairs they
 * covered, and clears the record. */
#define B200ORB_NSTAGE 7
int b200orb_batch_profile(b200orb_batch* b, int enable, int max_calls);
int b200orb_batch_profile_read(b200orb_batch* b, float* ms_per_stage, int* n_calls, long long* n_pairs);
/* kernel launches per stage of one b200orb_batch_run_device call */
int b200orb_batch_stage_launches(const b200orb_batch* b, int* launches_per_stage);

/* Host logic, no GPU needed: the chunk sizes b200orb_batch_run_host cuts a job of n_pairs into for an engine of max_pairs with
 * `lanes` compute lanes (1 or 2; run_host uses 2 unless B200ORB_HOST_LANES=1): half-capacity chunks with two lanes, a ramp
 * C/4, C/2 at both ends of a job of at least four chunks.  Writes at most `capacity` sizes (sizes may be NULL) and returns the
 * number of chunks (the reference has no counterpart: Tracking.py:95-112 builds one Frame per call). */
int b200orb_host_chunk_schedule(int max_pairs, int lanes, int n_pairs, int32_t* sizes, int capacity);

/* Host logic, no GPU needed: the FAST cell grid the engine plans for an H x W image -- the cells of ComputeKeyPointsOctTree's loops
 * (ORBextractor.cpp:770-806) in loop order, all levels.  Six ints per cell: level, iniX, iniY (window origin in level coordinates,
 * 3-px rim included), detection width and height (window minus the rim; 0 x 0 = a cell the reference skips or FAST cannot fill),
 * offset of the cell's candidate segment.  Writes at most `capacity` cells (cells may be NULL), returns the number of cells. */
int b200orb_plan_cells(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int H, int W, int32_t* cells,
                       int capacity);

/* ---------------------------------------------------------------------------------------------
 * BoW transform of the descriptors -- the tree descent of
 * TemplatedVocabulary.transform_feature (pyDBoW/TemplatedVocabulary.py:139-163) with FORB.distance
 * (pyDBoW/FORB.py:31-33), called through Frame.compute_BoW (Frame.py:123-125).
 * The vocabulary is given as arrays: node 0 is the root; the children of node i are child_ids[child_begin[i] ..
 * child_begin[i+1]) in the reference's child order; node_desc is uint8[n_nodes][32] (row 0 unused).
 * For every descriptor the transform returns the leaf it ends in (first strict minimum at every level) and the node
 * it passes at depth nid_level = L - levelsup (-1 if its path is shorter).  Word ids, weights, L1 normalisation and
 * the dictionaries are assembled by the caller exactly as the reference does (pyorbslam_b200/bow.py).
 * ------------------------------------------------------------------------------------------- */
typedef struct b200orb_vocab b200orb_vocab;
int b200orb_vocab_create(int n_nodes, const int32_t* child_begin, const int32_t* child_ids, const uint8_t* node_desc,
                         int device, b200orb_vocab** out);
void b200orb_vocab_destroy(b200orb_vocab* v);
/* descriptors from host memory: uint8[n][32] */
int b200orb_vocab_transform(b200orb_vocab* v, const uint8_t* desc, int n, int nid_level, int32_t* leaf_node, int32_t* level_node);
/* descriptors still resident on the device from the extractor's last b200orb_extract (no upload) */
int b200orb_vocab_transform_resident(b200orb_vocab* v, b200orb_extractor* e, int nid_level, int32_t* leaf_node,
                                     int32_t* level_node);

/* SURVEY.md 8(f) rank 1, first piece: all-pairs Hamming distances out[i*nB + j] = popcount(A[i] xor B[j]) between two sets of
 * 32-byte descriptors in host memory -- what ORBMatcher.descriptor_distance (ORBMatcher.py:12-14) computes one pair at a time
 * inside search_by_BoW_kf_f / search_by_BoW_kf_kf (ORBMatcher.py:21-213).  The greedy matching logic stays on the host
 * (pyorbslam_b200/matcher.py) and looks distances up in this matrix. */
int b200orb_hamming_matrix(int device, const uint8_t* A, int nA, const uint8_t* B, int nB, uint16_t* out);

/* SURVEY.md 8(f) rank 1, projection searches (ORBMatcher.search_by_projection_f_p, ORBMatcher.py:215-283, and
 * search_by_projection_f_f, ORBMatcher.py:291-393): Frame.get_features_in_area (Frame.py:373-416) for M queries at once, fused with
 * the Hamming distance from each query's descriptor to every feature it returns (ORBMatcher.descriptor_distance, ORBMatcher.py:12-14).
 *   kxy / koct / kdesc   the frame's N undistorted keypoints (x, y), octaves and 32-byte descriptors (mvKeysUn, mDescriptors)
 *   cell_start / cell_idx the frame's mGrid (Frame.assign_features_to_grid, Frame.py:153-159) as CSR over cell ix * rows + iy,
 *                         features of a cell in the order they were appended
 *   qxyr[M][3]           x, y, r of each query as the reference passes them; f32_mode != 0 when its scalar arithmetic ran in float32
 *                         (a NumPy float32 coordinate with Python-float radius, NumPy >= 2 promotion), 0 for float64
 *   qlvl[M][2]           min_level, max_level;   qcell[M][4]   n_min_cell_x, n_max_cell_x, n_min_cell_y, n_max_cell_y as evaluated by
 *                         the caller with the reference's own expressions (Frame.py:376-390; min > max = no cell)
 *   qdesc[M][32]         the query descriptors (MapPoint.get_descriptor())
 * Output (CSR over the queries): cand_start[M + 1], cand_idx / cand_dist[cap] in the order get_features_in_area returns them.
 * *total = number of candidates; B200ORB_E_RANGE with *total set when cap is too small (call again with larger buffers). */
int b200orb_area_hamming(int device, int f32_mode, int N, const float* kxy, const int32_t* koct, const uint8_t* kdesc, int cols, int rows,
                         const int32_t* cell_start, const int32_t* cell_idx, int M, const double* qxyr, const int32_t* qlvl,
                         const int32_t* qcell, const uint8_t* qdesc, int32_t* cand_start, int32_t* cand_idx, int32_t* cand_dist, int cap,
                         int32_t* total);
/* The order-dependent selection of the two searches on those candidate lists (host side: each decision depends on the matches made for
 * the earlier map points).  ok[c] = 0 marks candidates failing the right-image check (ORBMatcher.py:246-249, 345-349), occupied[j] != 0
 * features that hold a map point with observations() > 0 (updated in place), marks[q] != 0 map points with observations() > 0.
 * best[q] = matched feature or -1.  _ff: best distance <= th_high (ORBMatcher.py:335-366); _fp: best / second best with the octave
 * and ratio test (ORBMatcher.py:236-281). */
int b200orb_greedy_project_ff(int M, const int32_t* start, const int32_t* idx, const int32_t* dist, const uint8_t* ok, int N,
                              uint8_t* occupied, const uint8_t* marks, int th_high, int32_t* best);
int b200orb_greedy_project_fp(int M, const int32_t* start, const int32_t* idx, const int32_t* dist, const uint8_t* ok, int N,
                              uint8_t* occupied, const uint8_t* marks, const int32_t* koct, int th_high, double nnratio, int32_t* best);

/* pinned host memory helpers for callers without their own allocator */
int b200orb_host_alloc(void** p, size_t bytes);
int b200orb_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* B200ORB_H */
