// s4f_amg_setup.h -- levels of the GAMG hierarchy as the device set-up produces them (s4f_amg_setup.cu), fp64
#pragma once
#include <memory>
#include <vector>

#include "s4f_ctx.h"

struct AmgDevLevel {
    int n = 0, ld = 0, nSlices = 0; long long nnz = 0;
    DevBuf<int> slicePtr, col; DevBuf<double> a, dg;   // SELL-32 rows and the per-component diagonal [3*ld] (empty on level 0: the fine rows)
    DevBuf<int> parent;                                // [n]: cell -> cell of the next level (empty on the coarsest)
    DevBuf<int> childPtr, child;                       // children of this level's cells in the finer level, ascending (empty on level 0)
};

// mergeLevels pair-wise passes per solver level (OpenFOAM's GAMG keyword of the same name; 3 -> aggregates of up to 8 cells)
int s4f_amg_device_levels(s4fgpu_ctx* c, std::vector<std::unique_ptr<AmgDevLevel>>& levels, int coarsest, int mergeLevels);
