// Host-side procedural level generation (product code).  See levelgen.hpp.
//
// Reference behaviour reproduced here (file:line under /root/reference):
//   Env::reset reseed                         src/libs/env/src/env.cpp:61-62
//   TowerBuildingScenario::reset / platform   src/libs/scenarios/src/scenario_tower_building.cpp:19-78,129-154
//   EmptyPlatform::generate (floor + walls)   src/libs/scenarios/include/scenarios/platforms.hpp:165-188,325-329
//   VoxelGridComponent::addPlatform           src/libs/scenarios/include/scenarios/component_voxel_grid.hpp:73-106
//   greedy voxel -> box merge                 src/libs/scenarios/include/scenarios/component_voxel_grid.hpp:108-187
//   addBoundingBoxes / addTerrain             src/libs/scenarios/src/layout_utils.cpp:17-68
//   DefaultScenario::spawnAgents              src/libs/scenarios/include/scenarios/scenario_default.hpp:80-97
//   colour tables                             src/libs/env/include/env/const.hpp:25-143
#include "levelgen.hpp"

#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <unordered_map>
#include <unordered_set>

#include "hostmath.hpp"

namespace mv {

namespace {

// ---- colours -------------------------------------------------------------------------------------------------------
enum : uint32_t {
    C_YELLOW = 0xffdd3c, C_GREEN = 0x3bb372, C_LIGHT_GREEN = 0x50c878, C_BLUE = 0x2eb5d0, C_LIGHT_BLUE = 0xadd8e6, C_DARK_BLUE = 0x3a7fa6,
    C_DARK_NAVY = 0x2c3e50, C_ORANGE = 0xffb400, C_GREY = 0xb3b3b3, C_DARK_GREY = 0x555555, C_VERY_DARK_GREY = 0x222222, C_WHITE = 0xffffff,
    C_RED = 0xff0000, C_LIGHT_ORANGE = 0xffa770, C_VIOLET = 0xd468ee, C_LIGHT_PINK = 0xffe6e6, C_VL_YELLOW = 0xffffe6, C_VL_GREEN = 0xccffcc,
    C_VL_BLUE = 0xe6ecff, C_VL_GREY = 0xd9d9d9, C_VL_VIOLET = 0xf2e6ff, C_VL_ORANGE = 0xffebcc,
};
const uint32_t kPalette[22] = {C_YELLOW, C_GREEN, C_LIGHT_GREEN, C_BLUE, C_LIGHT_BLUE, C_DARK_BLUE, C_DARK_NAVY, C_ORANGE, C_GREY, C_DARK_GREY,
                               C_VERY_DARK_GREY, C_WHITE, C_RED, C_LIGHT_ORANGE, C_VIOLET, C_LIGHT_PINK, C_VL_YELLOW, C_VL_GREEN, C_VL_BLUE,
                               C_VL_GREY, C_VL_VIOLET, C_VL_ORANGE};
const uint32_t kLayoutColors[14] = {C_WHITE, C_VL_YELLOW, C_VL_GREEN, C_VL_BLUE, C_VL_GREY, C_VL_ORANGE, C_GREY, C_GREY, C_GREY, C_GREY,
                                    C_DARK_GREY, C_DARK_GREY, C_DARK_GREY, C_DARK_GREY};
int paletteIndex(uint32_t rgb) {
    for (int i = 0; i < 22; ++i)
        if (kPalette[i] == rgb) return i;
    return 0;
}

using Rng = std::mt19937;
int randRange(int low, int high, Rng &rng) { return std::uniform_int_distribution<>{low, high - 1}(rng); }
bool randomBool(Rng &rng) { return bool(randRange(0, 2, rng)); }
float frand(Rng &rng) { return std::uniform_real_distribution<float>{0, 1}(rng); }
uint32_t randomLayoutColor(Rng &rng) { return kLayoutColors[randRange(0, 14, rng)]; }

// ---- sparse voxel map with the reference's hash (voxel_grid.hpp:39-49): iteration order must match -------------------
struct I3 { int x, y, z; };
inline uint32_t voxKey(int x, int y, int z) { return uint32_t(((x + 512) << 20) + ((y + 512) << 10) + (z + 512)); }
inline I3 voxUnkey(uint32_t k) { return {int(k >> 20) - 512, int((k >> 10) & 1023) - 512, int(k & 1023) - 512}; }
struct KeyHash { size_t operator()(uint32_t k) const noexcept { return size_t(k); } };
struct Vox { uint8_t type = 0, terrain = 0; uint32_t color = C_WHITE; };
using VoxMap = std::unordered_map<uint32_t, Vox, KeyHash>;

struct IBox { int mn[3], mx[3]; };  // max exclusive for platform boxes, inclusive for merged boxes (as in the reference)

void fillBox(VoxMap &g, const IBox &b, uint8_t type, uint32_t color) {
    for (int x = b.mn[0]; x < b.mx[0]; ++x)
        for (int y = b.mn[1]; y < b.mx[1]; ++y)
            for (int z = b.mn[2]; z < b.mx[2]; ++z) {
                Vox v; v.type = type; v.terrain = 0; v.color = color;
                g[voxKey(x, y, z)] = v;
            }
}
void fillTerrain(VoxMap &g, const IBox &b, int terrain) {
    for (int x = b.mn[0]; x < b.mx[0]; ++x)
        for (int y = b.mn[1]; y < b.mx[1]; ++y)
            for (int z = b.mn[2]; z < b.mx[2]; ++z) {
                const uint32_t k = voxKey(x, y, z);
                if (!g.count(k)) g[k] = Vox();
                g.find(k)->second.terrain |= uint8_t(terrain);
            }
}

struct MergedGroup { uint8_t type; uint32_t color; std::vector<IBox> boxes; };

// greedy expansion in -x,+x,-y,+y,-z,+z order, seeds visited in the hash map's iteration order
std::vector<MergedGroup> mergeVoxels(const VoxMap &grid) {
    std::unordered_set<uint32_t, KeyHash> visited;
    const VoxMap snapshot = grid;  // the reference iterates a copy (component_voxel_grid.hpp:114)
    std::map<std::pair<uint8_t, uint32_t>, std::vector<IBox>> byType;
    std::vector<uint32_t> expansion;
    for (const auto &kv : snapshot) {
        const uint32_t key = kv.first;
        if (visited.count(key)) continue;
        visited.emplace(key);
        const uint8_t type = kv.second.type;
        const uint32_t color = kv.second.color;
        const I3 c = voxUnkey(key);
        int mn[3] = {c.x, c.y, c.z}, mx[3] = {c.x, c.y, c.z};
        for (int axis = 0; axis < 3; ++axis)
            for (int sign = -1; sign <= 1; sign += 2) {
                while (true) {
                    int lo[3], hi[3];
                    for (int a = 0; a < 3; ++a) {
                        if (a == axis) lo[a] = hi[a] = sign > 0 ? mx[a] + 1 : mn[a] - 1;
                        else { lo[a] = mn[a]; hi[a] = mx[a]; }
                    }
                    expansion.clear();
                    bool ok = true;
                    for (int x = lo[0]; x <= hi[0] && ok; ++x)
                        for (int y = lo[1]; y <= hi[1] && ok; ++y)
                            for (int z = lo[2]; z <= hi[2]; ++z) {
                                const uint32_t k = voxKey(x, y, z);
                                const auto it = grid.find(k);
                                if (it == grid.end() || it->second.type != type || it->second.color != color || visited.count(k)) { ok = false; break; }
                                expansion.push_back(k);
                            }
                    if (!ok) break;
                    for (uint32_t k : expansion) visited.emplace(k);
                    if (sign > 0) mx[axis] += 1; else mn[axis] -= 1;
                }
            }
        IBox b;
        for (int a = 0; a < 3; ++a) { b.mn[a] = mn[a]; b.mx[a] = mx[a]; }
        byType[{type, color}].push_back(b);
    }
    std::vector<MergedGroup> out;
    for (auto &kv : byType) out.push_back({kv.first.first, kv.first.second, std::move(kv.second)});
    return out;
}

void setTerrainModel(MvTerrain &t, const IBox &bb, uint32_t color) {
    using namespace mvh;
    const float sx = float(bb.mx[0] - bb.mn[0]) * 1.0f, sy = 1.0f * 1.0f, sz = float(bb.mx[2] - bb.mn[2]) * 1.0f;
    const float px = bb.mn[0] * 1.0f + sx / 2, py = bb.mn[1] * 1.0f, pz = bb.mn[2] * 1.0f + sz / 2;
    M4 m = mul(scaling(0.5f, 0.025f, 0.5f), identity());
    m = mul(scaling(sx, sy, sz), m);
    m = mul(translation(0.0f, 0.025f, 0.0f), m);
    m = mul(translation(px, py, pz), m);
    std::memcpy(t.model, &m.c[0][0], 64);
    t.color = paletteIndex(color);
}

}  // namespace

int scenarioFromName(const std::string &name) {
    std::string n;
    for (char ch : name) n.push_back(char(std::tolower(ch)));
    if (n == "towerbuilding") return MV_SCENARIO_TOWER;
    return -1;
}

FloatParams defaultFloatParams(int) {  // scenario.hpp:225-231
    return {{"episodeLengthSec", 60.0f}, {"verticalLookLimitRad", 0.2f}, {"useUIRewardIndicators", 0.0f}};
}

std::vector<std::pair<std::string, float>> defaultRewardShaping(int scenario) {
    if (scenario == MV_SCENARIO_TOWER)  // scenario_tower_building.hpp:44-52
        return {{"teamSpirit", 0.1f}, {"towerPickedUpObject", 0.1f}, {"towerVisitedBuildingZoneWithObject", 0.1f}, {"towerBuildingReward", 1.0f}};
    return {{"teamSpirit", 0.0f}};
}

int rewardSlot(int scenario, const std::string &key) {
    if (key == "teamSpirit") return MV_R_TEAM_SPIRIT;
    if (scenario == MV_SCENARIO_TOWER) {
        if (key == "towerPickedUpObject") return MV_R_TOWER_PICKED_UP;
        if (key == "towerVisitedBuildingZoneWithObject") return MV_R_TOWER_VISITED_BZ;
        if (key == "towerBuildingReward") return MV_R_TOWER_BUILDING;
    }
    return -1;
}

LevelGenerator::LevelGenerator(int scenario, int numAgents, const FloatParams &params) : scenario_(scenario), numAgents_(numAgents), params_(params) {}

void LevelGenerator::generate(LevelOut &out, int serial, int gridCells) {
    std::memset(&out.level, 0, sizeof(MvLevel));
    // Env::reset: reseed the env stream from itself
    const auto sd = randRange(0, 1 << 30, rng_);
    rng_.seed((unsigned long)sd);
    out.level.serial = serial;
    out.level.scenario = scenario_;
    out.level.episode_len_base = params_.at("episodeLengthSec");
    out.level.look_limit = params_.at("verticalLookLimitRad");
    switch (scenario_) {
        case MV_SCENARIO_TOWER: generateTower(out); break;
        default: throw std::runtime_error("unsupported scenario");
    }
    const MvLevel &L = out.level;
    if (L.grid_dim[0] * L.grid_dim[1] * L.grid_dim[2] > gridCells) throw std::runtime_error("level exceeds the dense grid capacity");
}

void LevelGenerator::generateTower(LevelOut &out) {
    MvLevel &L = out.level;
    Rng &rng = rng_;
    const int A = numAgents_;

    uint32_t layoutColor = randomLayoutColor(rng);
    while (layoutColor == C_DARK_GREY) layoutColor = randomLayoutColor(rng);  // != BUILDING_ZONE

    // TowerBuildingPlatform::init
    const int height = randRange(5, 7, rng);
    int length = randRange(12, 30, rng);
    int width = randRange(12, 25, rng);
    const int bzL = randRange(3, 9, rng), bzW = randRange(3, 9, rng);
    const int matL = randRange(2, 8, rng), matW = randRange(2, 8, rng);
    length = std::max(bzL + matL + 3, length);
    width = std::max(bzW + matW + 3, width);
    const int bzX = randRange(1, length - bzL - 1, rng);
    const int bzZ = randRange(1, width - bzW - 1, rng);
    const int matX = randRange(1, length - matL - 1, rng);
    const int matZ = randRange(1, width - matW - 1, rng);

    std::vector<I3> cand;
    for (int x = 1; x < length - 1; ++x)
        for (int z = 1; z < width - 1; ++z) cand.push_back({x, 2, z});
    std::shuffle(cand.begin(), cand.end(), rng);

    std::vector<I3> agentSpawn(cand.begin(), cand.begin() + std::min(A, int(cand.size())));
    const int spawnIdx = int(agentSpawn.size());
    const int maxRandomObjects = std::min(int(cand.size()) - A, 25);
    const int spawnObjects = randRange(0, std::max(1, maxRandomObjects), rng);
    std::vector<I3> objs(cand.begin() + spawnIdx, cand.begin() + spawnIdx + spawnObjects);
    for (auto &c : objs) {
        if (c.x >= matX && c.x < matX + matL && c.z >= matZ && c.z < matZ + matW) continue;
        c.y -= 1;
    }
    for (int x = matX; x < matX + matL; ++x)
        for (int z = matZ; z < matZ + matW; ++z) objs.push_back({x, 1, z});
    while (int(agentSpawn.size()) < A) agentSpawn.push_back(agentSpawn[0]);

    // generate(): floor, walls S N E W, building-zone terrain box; then (GCC argument order) drawWalls, wall colour
    const bool drawWalls = randomBool(rng);
    const uint32_t wallColor = randomLayoutColor(rng);

    VoxMap grid{100};
    fillBox(grid, {{0, 0, 0}, {length, 1, width}}, MV_SOLID | MV_OPAQUE, layoutColor);
    const uint8_t wallType = uint8_t(MV_SOLID | (drawWalls ? MV_OPAQUE : 0));
    fillBox(grid, {{0, 0, 0}, {1, height, width}}, wallType, wallColor);
    fillBox(grid, {{length - 1, 0, 0}, {length, height, width}}, wallType, wallColor);
    fillBox(grid, {{0, 0, 0}, {length, height, 1}}, wallType, wallColor);
    fillBox(grid, {{0, 0, width - 1}, {length, height, width}}, wallType, wallColor);
    const IBox bz{{bzX, 1, bzZ}, {bzX + bzL, 1, bzZ + bzW}};
    fillTerrain(grid, bz, 4);  // empty y-range: sets nothing, exactly like the reference

    // DefaultScenario::spawnAgents
    for (int i = 0; i < A; ++i) {
        const float yaw = frand(rng) * 3.14159265358979323846f * 2;
        mvh::yawBasis(yaw, L.spawn_basis[i]);
        const float sx = float(agentSpawn[i].x) + 0.5f, sy = float(agentSpawn[i].y) + 0.0f, sz = float(agentSpawn[i].z) + 0.5f;
        L.spawn_pos[i][0] = sx; L.spawn_pos[i][1] = sy + 1.75f; L.spawn_pos[i][2] = sz;
        L.init_pos[i][0] = float(agentSpawn[i].x); L.init_pos[i][1] = float(agentSpawn[i].y); L.init_pos[i][2] = float(agentSpawn[i].z);
    }

    // addEpisodeDrawables: merged boxes (skip VOXEL_EMPTY groups), terrain slabs, movable objects
    int ns = 0;
    for (const auto &g : mergeVoxels(grid)) {
        if (g.type == 0) continue;
        for (const auto &b : g.boxes) {
            if (ns >= MV_MAX_STATIC) throw std::runtime_error("too many static boxes");
            MvBox &sb = L.statics[ns++];
            for (int a = 0; a < 3; ++a) {
                sb.h[a] = (float(b.mx[a] - b.mn[a] + 1) / 2) * 1.0f;
                sb.c[a] = (float(b.mn[a] + b.mx[a]) / 2 + 0.5f) * 1.0f;
            }
            sb.flags = g.type;
            sb.color = paletteIndex(g.color);
        }
    }
    L.n_static = ns;
    L.n_terrain = 0;
    if (bz.mx[0] - bz.mn[0] > 0) setTerrainModel(L.terrain[L.n_terrain++], bz, C_DARK_GREY);
    if (int(objs.size()) > MV_MAX_OBJECTS - 1) throw std::runtime_error("too many movable objects");
    L.n_obj = int(objs.size());
    L.n_movable = int(objs.size());
    for (int i = 0; i < L.n_obj; ++i) {
        L.obj_voxel[i][0] = int16_t(objs[i].x); L.obj_voxel[i][1] = int16_t(objs[i].y); L.obj_voxel[i][2] = int16_t(objs[i].z);
        L.obj_voxel[i][3] = int16_t(paletteIndex(C_LIGHT_BLUE));
    }
    for (int a = 0; a < 3; ++a) { L.bz_min[a] = bz.mn[a]; L.bz_max[a] = bz.mx[a]; }

    // dense grid: the room's bounding box, with head-room above the walls for stacked objects
    L.grid_org[0] = 0; L.grid_org[1] = 0; L.grid_org[2] = 0;
    L.grid_dim[0] = length; L.grid_dim[1] = height + 18; L.grid_dim[2] = width;
    const int cells = L.grid_dim[0] * L.grid_dim[1] * L.grid_dim[2];
    out.solid.assign(size_t(cells + 31) / 32, 0u);
    for (const auto &kv : grid) {
        if (!(kv.second.type & MV_SOLID)) continue;
        const I3 c = voxUnkey(kv.first);
        const int idx = ((c.x - L.grid_org[0]) * L.grid_dim[1] + (c.y - L.grid_org[1])) * L.grid_dim[2] + (c.z - L.grid_org[2]);
        out.solid[size_t(idx >> 5)] |= 1u << (idx & 31);
    }
}

}  // namespace mv
