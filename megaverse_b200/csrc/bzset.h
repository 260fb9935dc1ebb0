// Emulation of libstdc++'s std::unordered_set<VoxelCoords> ITERATION ORDER for TowerBuilding's objectsInBuildingZone
// (src/libs/scenarios/include/scenarios/scenario_tower_building.hpp:78, summed in float by calculateTowerReward,
// src/libs/scenarios/src/scenario_tower_building.cpp:232-241) -- the order of that float sum decides the reward bits.
//
// libstdc++'s _Hashtable keeps ONE singly linked list; nodes of a bucket are contiguous.  Inserting a node whose bucket
// is empty puts it at the list head, otherwise at the head of its bucket's run (_M_insert_bucket_begin); a rehash
// re-inserts every node in list order with the same rule (_M_rehash_aux, unique keys); clear() keeps the bucket count and
// the policy's next-resize threshold.  With max_load_factor 1 the bucket count walks 1 -> 13 -> 29 -> 59 -> 127 -> 257
// (_Prime_rehash_policy::_M_need_rehash / _M_next_bkt).  Validated against a real std::unordered_set in tests/.
#pragma once
#include "mv_types.h"

#ifdef __CUDACC__
#define MV_HD __host__ __device__ __forceinline__
#else
#define MV_HD inline
#endif

MV_HD uint32_t mvVoxelHash(int x, int y, int z) { return uint32_t(((x + 512) << 20) + ((y + 512) << 10) + (z + 512)); }  // voxel_grid.hpp:39-49

MV_HD void mvBzInit(MvEnvState &s) { s.bz_count = 0; s.bz_nb = 1; s.bz_next_resize = 0; }
MV_HD void mvBzClear(MvEnvState &s) { s.bz_count = 0; }

MV_HD int mvBzFind(const MvEnvState &s, int x, int y, int z) {
    for (int i = 0; i < s.bz_count; ++i)
        if (s.bz_items[i][0] == x && s.bz_items[i][1] == y && s.bz_items[i][2] == z) return i;
    return -1;
}

MV_HD void mvBzInsertRaw(int16_t (*items)[4], int &count, int nb, int x, int y, int z) {
    const uint32_t b = mvVoxelHash(x, y, z) % uint32_t(nb);
    int at = 0;
    for (int i = 0; i < count; ++i)
        if (mvVoxelHash(items[i][0], items[i][1], items[i][2]) % uint32_t(nb) == b) { at = i; break; }
    for (int i = count; i > at; --i) {
        items[i][0] = items[i - 1][0]; items[i][1] = items[i - 1][1]; items[i][2] = items[i - 1][2];
    }
    items[at][0] = int16_t(x); items[at][1] = int16_t(y); items[at][2] = int16_t(z);
    ++count;
}

MV_HD void mvBzInsert(MvEnvState &s, int x, int y, int z) {
    if (mvBzFind(s, x, y, z) >= 0) return;
    if (s.bz_count >= MV_MAX_OBJECTS - 1) return;
    if (s.bz_count + 1 > s.bz_next_resize) {
        const int nb = s.bz_nb == 1 ? 13 : (s.bz_nb == 13 ? 29 : (s.bz_nb == 29 ? 59 : (s.bz_nb == 59 ? 127 : 257)));
        // rehash: re-insert in list order under the new bucket count.  Done in place back to front is not order
        // preserving, so walk a copy of the (short) list.
        int16_t tmp[MV_MAX_OBJECTS][4];
        const int n = s.bz_count;
        for (int i = 0; i < n; ++i) { tmp[i][0] = s.bz_items[i][0]; tmp[i][1] = s.bz_items[i][1]; tmp[i][2] = s.bz_items[i][2]; }
        int cnt = 0;
        for (int i = 0; i < n; ++i) mvBzInsertRaw(s.bz_items, cnt, nb, tmp[i][0], tmp[i][1], tmp[i][2]);
        s.bz_nb = nb;
        s.bz_next_resize = nb;
    }
    mvBzInsertRaw(s.bz_items, s.bz_count, s.bz_nb, x, y, z);
}

MV_HD void mvBzErase(MvEnvState &s, int x, int y, int z) {
    const int at = mvBzFind(s, x, y, z);
    if (at < 0) return;
    for (int i = at; i + 1 < s.bz_count; ++i) {
        s.bz_items[i][0] = s.bz_items[i + 1][0]; s.bz_items[i][1] = s.bz_items[i + 1][1]; s.bz_items[i][2] = s.bz_items[i + 1][2];
    }
    --s.bz_count;
}
