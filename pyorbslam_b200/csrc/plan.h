// Geometry plan shared by host code and kernels (passed by value as a __grid_constant__ kernel parameter).
// One Plan describes how ONE image of size H x W is laid out in the per-slot workspace; a batch is S slots.
#pragma once
#include <stdint.h>

#define ORB_MAX_LEVELS 16
#define ORB_EDGE 19          // EDGE_THRESHOLD, ORBextractor.cpp:74
#define ORB_DET_ORIGIN 16    // EDGE_THRESHOLD - 3: origin of the FAST detection area, ORBextractor.cpp:772

struct LevelGeom {
    int w, h;               // level image (ROI) size, ORBextractor.cpp:1110-1111
    int pitch, rows;        // bordered buffer: physical pitch (multiple of 64 B) and rows = h + 38
    int pyr_ofs;            // byte offset of the bordered buffer inside the slot's pyramid blob
    int blur_pitch, blur_ofs;
    int maxBX, maxBY;       // w - 16, h - 16  (maxBorderX/Y, ORBextractor.cpp:774-775)
    int nCols, nRows, wCell, hCell;   // cell grid, ORBextractor.cpp:783-786 (nRows == 0: level has no cells)
    int cell_ofs;           // first cell of this level in the slot's cell-count array
    int cell_cap;           // candidate capacity of one cell
    int cand_ofs;           // u32 index of the level's candidate storage in the slot's candidate blob
    int blur_cta_ofs, blur_tiles_x;
    int quota;              // mnFeaturesPerLevel
    int kp_ofs, kp_cap;     // level segment in the slot's level-keypoint array
    int nIni;               // DistributeOctTree root count, ORBextractor.cpp:543
    float hX;               // root width, ORBextractor.cpp:545
    int root_ofs;           // offset of this level's root_of[x] table (k_octree): (int)((float)x / hX) for x = 0 .. maxBX - 16
    float sf, isf;          // mvScaleFactor / mvInvScaleFactor
    int psize;              // (int)(31 * sf), ORBextractor.cpp:834
    int xtab_ofs, ytab_ofs; // resize tables (level >= 1)
};

struct Plan {
    int nlevels, H, W, iniTh, minTh;
    int pyr_bytes, blur_bytes;   // per slot
    int ncells, cand_entries;    // per slot
    int kp_total;                // per slot: sum of kp_cap == row capacity of the output arrays
    int blur_ctas;
    int max_cells_level;         // largest cell count of one level
    int umax[16];
    LevelGeom lv[ORB_MAX_LEVELS];
};

// resize lookup tables (built on the host with the exact OpenCV float arithmetic)
struct XTab { int sx; short a0, a1; };            // source column + horizontal 11-bit coefficients
struct YTab { short y0, y1, b0, b1; };            // clamped source rows + vertical coefficients
// One entry per group of 4 consecutive bordered output columns (= one k_resize thread): where its 12-byte source window starts
// (word index in a bordered source row), how far column 0's left sample sits inside it, the byte-pair selectors of the 4
// columns relative to column 0 and their packed coefficients.  shift8 == 0xffffffff: the group is not an interior one
// (reflected border columns, or columns more than 6 source bytes apart) and takes the per-byte path through XTab.
#define MOM_STEPS 11   // IC_Angle: 31 disc rows, 3 per step
struct XGroup { int wofs; unsigned shift8, sel01, sel23, cf[4]; };

// pyramid addressing for the stereo SAD stage (device-resident pyramids or uploaded GetImagePyramid() views)
struct StereoGeom {
    int nlevels;
    float sf[ORB_MAX_LEVELS], isf[ORB_MAX_LEVELS];
    int w[ORB_MAX_LEVELS], h[ORB_MAX_LEVELS];
    int plog[ORB_MAX_LEVELS];    // logical pitch of the buffer the view is cut from (w+38 resident, w uploaded)
    unsigned magic[ORB_MAX_LEVELS];   // ceil(2^32 / plog): division by plog as multiply-high + one correction
    int pitch[ORB_MAX_LEVELS];   // physical pitch
    int off0[ORB_MAX_LEVELS];    // logical linear offset of view element (0,0): 19*plog+19 resident, 0 uploaded
    int vstride[ORB_MAX_LEVELS]; // logical offset between view rows: w for the reference's sheared caster view, plog for the true image
    long long base[ORB_MAX_LEVELS];  // byte offset of the level inside one image's blob
    int nRows;                   // rows of level 0 (Frame.py:167)
};
