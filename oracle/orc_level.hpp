// ORACLE (test infrastructure, NOT product code): level-layout pieces, restated from
//   src/libs/util/include/util/voxel_grid.hpp:14-165      (VoxelCoords, hash, VoxelGrid)
//   src/libs/util/include/util/util.hpp:25-49              (Rng, randRange, randomBool, frand, randomSample)
//   src/libs/env/include/env/voxel_state.hpp:10-45         (VoxelType, VoxelState)
//   src/libs/env/include/env/const.hpp:25-143              (ColorRgb, palettes)
//   src/libs/scenarios/include/scenarios/platforms.hpp:19-330 (BoundingBox, MagnumAABB, Platform, EmptyPlatform)
//   src/libs/scenarios/include/scenarios/component_voxel_grid.hpp:17-187 (VoxelGridComponent, toBoundingBoxes)
// Written in the reference's own style on purpose (hash maps, a float scene-graph transform per
// platform) so that it is an independent restatement from the product's flat level generator.
#pragma once
#include <algorithm>
#include <map>
#include <memory>
#include <random>
#include <set>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "orc_math.hpp"

namespace orc {

struct VoxelCoords {
    int x = 0, y = 0, z = 0;
    VoxelCoords() = default;
    VoxelCoords(int x_, int y_, int z_) : x(x_), y(y_), z(z_) {}
    bool operator==(const VoxelCoords &o) const { return x == o.x && y == o.y && z == o.z; }
    bool operator!=(const VoxelCoords &o) const { return !(*this == o); }
};

struct VoxelHash {  // voxel_grid.hpp:39-49
    std::size_t operator()(const VoxelCoords &v) const noexcept {
        const int x = v.x + 512, y = v.y + 512, z = v.z + 512;
        return std::size_t((x << 20) + (y << 10) + z);
    }
};

// voxel_grid.hpp:18-21  lround(floor(v))
inline VoxelCoords toVoxel(Vec3 v) { return {int(lroundf(floorf(v.x))), int(lroundf(floorf(v.y))), int(lroundf(floorf(v.z)))}; }

using Rng = std::mt19937;
inline int randRange(int low, int high, Rng &rng) { return std::uniform_int_distribution<>{low, high - 1}(rng); }
inline bool randomBool(Rng &rng) { return bool(randRange(0, 2, rng)); }
inline float frand(Rng &rng) { return std::uniform_real_distribution<float>{0, 1}(rng); }
template <typename C> auto randomSample(const C &c, Rng &rng) { const auto idx = randRange(0, int(c.size()), rng); return c[idx]; }

enum ColorRgb : uint32_t {
    YELLOW = 0xffdd3c, GREEN = 0x3bb372, LIGHT_GREEN = 0x50c878, BLUE = 0x2eb5d0, LIGHT_BLUE = 0xadd8e6,
    DARK_BLUE = 0x3a7fa6, DARK_NAVY = 0x2c3e50, ORANGE = 0xffb400, GREY = 0xb3b3b3, DARK_GREY = 0x555555,
    VERY_DARK_GREY = 0x222222, WHITE = 0xffffff, RED = 0xff0000, LIGHT_ORANGE = 0xffa770, VIOLET = 0xd468ee,
    LIGHT_PINK = 0xffe6e6, VERY_LIGHT_YELLOW = 0xffffe6, VERY_LIGHT_GREEN = 0xccffcc, VERY_LIGHT_BLUE = 0xe6ecff,
    VERY_LIGHT_GREY = 0xd9d9d9, VERY_LIGHT_VIOLET = 0xf2e6ff, VERY_LIGHT_ORANGE = 0xffebcc,
    LAYOUT_DEFAULT = WHITE, AGENT_EYES = DARK_NAVY, MOVABLE_BOX = LIGHT_BLUE, EXIT_PAD = LIGHT_GREEN, BUILDING_ZONE = DARK_GREY,
};
static const ColorRgb allColors[] = {YELLOW, GREEN, LIGHT_GREEN, BLUE, LIGHT_BLUE, DARK_BLUE, DARK_NAVY, ORANGE, GREY, DARK_GREY,
    VERY_DARK_GREY, WHITE, RED, LIGHT_ORANGE, VIOLET, LIGHT_PINK, VERY_LIGHT_YELLOW, VERY_LIGHT_GREEN, VERY_LIGHT_BLUE,
    VERY_LIGHT_GREY, VERY_LIGHT_VIOLET, VERY_LIGHT_ORANGE};
static const int numColors = 22;
static const ColorRgb agentColors[] = {YELLOW, GREEN, BLUE, ORANGE, VIOLET, VERY_DARK_GREY, RED};
static const int numAgentColors = 7;
static const ColorRgb objectColors[] = {YELLOW, GREEN, LIGHT_GREEN, BLUE, LIGHT_BLUE, DARK_BLUE, ORANGE, GREY, DARK_GREY, WHITE, RED,
    LIGHT_ORANGE, VIOLET, LIGHT_PINK};
static const int numObjectColors = 14;
static const ColorRgb layoutColors[] = {LAYOUT_DEFAULT, VERY_LIGHT_YELLOW, VERY_LIGHT_GREEN, VERY_LIGHT_BLUE, VERY_LIGHT_GREY,
    VERY_LIGHT_ORANGE, GREY, GREY, GREY, GREY, DARK_GREY, DARK_GREY, DARK_GREY, DARK_GREY};
static const int numLayoutColors = 14;
inline ColorRgb sampleRandomColor(Rng &rng) { return allColors[randRange(0, numColors, rng)]; }
inline ColorRgb randomObjectColor(Rng &rng) { return objectColors[randRange(0, numObjectColors, rng)]; }
inline ColorRgb randomLayoutColor(Rng &rng) { return layoutColors[randRange(0, numLayoutColors, rng)]; }
inline int paletteIndex(ColorRgb c) { for (int i = 0; i < numColors; ++i) if (allColors[i] == c) return i; return -1; }

enum VoxelType { VOXEL_EMPTY = 0, VOXEL_SOLID = 1, VOXEL_OPAQUE = 2 };
enum TerrainType { TERRAIN_NONE = 0, TERRAIN_EXIT = 1, TERRAIN_LAVA = 2, TERRAIN_BUILDING_ZONE = 4 };
enum { WALLS_SOUTH = 1, WALLS_NORTH = 2, WALLS_WEST = 4, WALLS_EAST = 8, WALLS_NONE = 0, WALLS_ALL = 15 };
inline ColorRgb terrainColor(int t) { return t == TERRAIN_EXIT ? EXIT_PAD : t == TERRAIN_LAVA ? RED : BUILDING_ZONE; }

struct Voxel {  // VoxelState + VoxelWithPhysicsObjects (+ rewardObject for Collect)
    uint8_t voxelType = VOXEL_EMPTY, terrain = 0;
    ColorRgb color = LAYOUT_DEFAULT;
    int physicsObject = -1;  // index into Env::objects, -1 == nullptr
    int rewardObject = -1;   // Collect: index of the reward drawable
    int reward = 0;
    bool solid() const { return voxelType & VOXEL_SOLID; }
    bool empty() const { return !solid(); }
};

class VoxelGrid {  // voxel_grid.hpp:57-165
public:
    using HashMap = std::unordered_map<VoxelCoords, Voxel, VoxelHash>;
    explicit VoxelGrid(size_t voxelCount = 100, Vec3 origin = {0, 0, 0}, float voxelSize = 1) : voxelCount(voxelCount), grid(voxelCount), origin(origin), voxelSize(voxelSize) {}
    void clear() { grid = HashMap{voxelCount}; }
    bool hasVoxel(const VoxelCoords &c) const { return bool(grid.count(c)); }
    Voxel *get(const VoxelCoords &c) { auto it = grid.find(c); return it == grid.end() ? nullptr : &it->second; }
    const Voxel *get(const VoxelCoords &c) const { auto it = grid.find(c); return it == grid.end() ? nullptr : &it->second; }
    Voxel *getWithVector(Vec3 v) { return get(getCoords(v)); }
    void set(const VoxelCoords &c, const Voxel &v) { grid[c] = v; }
    void remove(const VoxelCoords &c) { grid.erase(c); }
    VoxelCoords getCoords(Vec3 v) const {
        Vec3 f{(v.x - origin.x) / voxelSize, (v.y - origin.y) / voxelSize, (v.z - origin.z) / voxelSize};
        return toVoxel(f);
    }
    const HashMap &getHashMap() const { return grid; }
    float getVoxelSize() const { return voxelSize; }
    size_t voxelCount;
    HashMap grid;
    Vec3 origin;
    float voxelSize;
};

struct BoundingBox {  // platforms.hpp:60-110
    VoxelCoords min, max;
    BoundingBox() = default;
    BoundingBox(VoxelCoords a, VoxelCoords b) : min(a), max(b) {}
    BoundingBox(int a, int b, int c, int d, int e, int f) : min(a, b, c), max(d, e, f) {}
    void addPoint(const VoxelCoords &v) {
        if (v.x < min.x) min.x = v.x;
        if (v.x > max.x) max.x = v.x;
        if (v.y < min.y) min.y = v.y;
        if (v.y > max.y) max.y = v.y;
        if (v.z < min.z) min.z = v.z;
        if (v.z > max.z) max.z = v.z;
    }
    void sort() {
        if (min.x > max.x) std::swap(min.x, max.x);
        if (min.y > max.y) std::swap(min.y, max.y);
        if (min.z > max.z) std::swap(min.z, max.z);
    }
    bool collidesWith(const BoundingBox &o) const {
        if (max.x <= o.min.x) return false;
        if (min.x >= o.max.x) return false;
        if (max.y <= o.min.y) return false;
        if (min.y >= o.max.y) return false;
        if (max.z <= o.min.z) return false;
        if (min.z >= o.max.z) return false;
        return true;
    }
};

// A scene-graph node reduced to what the platform code needs: parent pointer + local matrix.
struct Node {
    Node *parent = nullptr;
    Mat4 local = mat4Identity();
    Mat4 absolute() const { return parent ? mul(parent->absolute(), local) : local; }  // Object::absoluteTransformation
    void translateLocal(Vec3 t) { local = mul(local, mat4Translation(t)); }
    void rotateYLocal(float rad) { local = mul(local, mat4RotationY(rad)); }
};

struct MagnumAABB {  // platforms.hpp:113-134
    Node *min, *max;
    BoundingBox boundingBox() const {
        BoundingBox bb;
        Vec3 a = translationOf(min->absolute()), b = translationOf(max->absolute());
        bb.min = {int(lroundf(a.x)), int(lroundf(a.y)), int(lroundf(a.z))};
        bb.max = {int(lroundf(b.x)), int(lroundf(b.y)), int(lroundf(b.z))};
        bb.sort();
        return bb;
    }
};

using FloatParams = std::map<std::string, float>;

class Platform {  // platforms.hpp:137-300
public:
    Platform(Node *parent, Rng &rng, int walls, const FloatParams &params) : rng(rng), walls(walls), params(params) {
        root = newNode(parent);
    }
    virtual ~Platform() = default;
    virtual void init() = 0;
    virtual void generate() = 0;

    Node *newNode(Node *parent) { nodes.emplace_back(std::make_unique<Node>()); nodes.back()->parent = parent; return nodes.back().get(); }
    MagnumAABB makeAABB(const BoundingBox &bb) {
        MagnumAABB a{newNode(root), newNode(root)};
        a.min->translateLocal({float(bb.min.x), float(bb.min.y), float(bb.min.z)});
        a.max->translateLocal({float(bb.max.x), float(bb.max.y), float(bb.max.z)});
        return a;
    }
    // degrees(90.0) -> Rad: T(value)*pi/T(180)  (Magnum Math/Angle.h)
    virtual void rotateCCW(int) { root->rotateYLocal(90.0f * 3.14159265358979323846f / 180.0f); root->translateLocal({-1, 0, -1}); }
    virtual void rotateCW(int previousPlatformWidth) {
        root->rotateYLocal(-90.0f * 3.14159265358979323846f / 180.0f);
        root->translateLocal({float(previousPlatformWidth) - 1, 0, -float(width) + 1});
    }
    virtual void addFloor() {
        layoutBoxes.emplace_back(makeAABB({0, 0, 0, length, 1, width}));
        nextPlatformAnchor = newNode(root);
        nextPlatformAnchor->translateLocal({float(length), 0, 0});
        boundingBoxDirty = true;
    }
    virtual void addWalls() {
        if (walls & WALLS_SOUTH) wallBoxes.emplace_back(makeAABB({0, 0, 0, 1, height, width}));
        if (walls & WALLS_NORTH) wallBoxes.emplace_back(makeAABB({length - 1, 0, 0, length, height, width}));
        if (walls & WALLS_EAST) wallBoxes.emplace_back(makeAABB({0, 0, 0, length, height, 1}));
        if (walls & WALLS_WEST) wallBoxes.emplace_back(makeAABB({0, 0, width - 1, length, height, width}));
        boundingBoxDirty = true;
    }
    virtual BoundingBox platformBoundingBox() {
        if (!boundingBoxDirty) return outerBoundingBox;
        if (!layoutBoxes.empty()) outerBoundingBox = layoutBoxes.front().boundingBox();
        else if (!wallBoxes.empty()) outerBoundingBox = wallBoxes.front().boundingBox();
        else outerBoundingBox = BoundingBox{};
        boundingBoxDirty = false;
        for (auto *boxes : {&layoutBoxes, &wallBoxes})
            for (auto &box : *boxes) {
                const auto bb = box.boundingBox();
                outerBoundingBox.addPoint(bb.min);
                outerBoundingBox.addPoint(bb.max);
            }
        return outerBoundingBox;
    }
    virtual bool collidesWith(Platform &other) { return platformBoundingBox().collidesWith(other.platformBoundingBox()); }
    virtual std::vector<Vec3> agentSpawnPoints(int numAgents) {
        std::vector<Vec3> spawnPoints;
        std::set<std::pair<int, int>> used;
        for (int i = 0; i < numAgents; ++i) {
            int x, y, z;
            for (int attempt = 0; attempt < 10; ++attempt) {
                x = randRange(1, length - 1, rng);
                z = randRange(1, width - 1, rng);
                if (used.count({x, z})) continue;
                y = occupancy[{x, z}] + 1;
                occupancy[{x, z}] += 2;
                spawnPoints.emplace_back(float(x), float(y), float(z));
                used.emplace(x, z);
                break;
            }
        }
        return spawnPoints;
    }
    virtual int requiresMovableBoxesToTraverse() { return 0; }
    virtual std::vector<VoxelCoords> generateObjectPositions(int n) {
        std::vector<VoxelCoords> boxes;
        constexpr int maxAttempts = 10;
        for (int i = 0; i < n; ++i)
            for (int attempt = 0; attempt < 10; ++attempt) {
                const int x = randRange(1, length - 1, rng);
                const int z = randRange(1, width - 1, rng);
                if (occupancy[{x, z}] < 2 || attempt >= maxAttempts - 1) {
                    const int y = ++occupancy[{x, z}];
                    boxes.emplace_back(x, y, z);
                    break;
                }
            }
        return adjustTransformation(boxes);
    }
    std::vector<VoxelCoords> adjustTransformation(std::vector<VoxelCoords> &coords) const {
        const Mat4 abs = root->absolute();
        for (auto &c : coords) c = toVoxel(transformPoint(abs, {c.x + 0.5f, c.y + 0.5f, c.z + 0.5f}));
        return coords;
    }
    int param(const std::string &s) const { return int(lroundf(params.at(s))); }
    virtual bool isMaxDifficulty() const { return false; }

    Rng &rng;
    int walls{};
    int length{}, height{}, width{};
    std::vector<MagnumAABB> layoutBoxes, wallBoxes;
    std::map<int, std::vector<MagnumAABB>> terrainBoxes;
    Node *nextPlatformAnchor{}, *root{};
    bool boundingBoxDirty = true;
    BoundingBox outerBoundingBox;
    std::map<std::pair<int, int>, int> occupancy;
    const FloatParams &params;
    std::vector<std::unique_ptr<Node>> nodes;
};

class EmptyPlatform : public Platform {  // platforms.hpp:302-330
public:
    EmptyPlatform(Node *parent, Rng &rng, int walls, const FloatParams &params, int w = -1) : Platform(parent, rng, walls, params) { width = w; }
    void init() override {
        length = randRange(4, 10, rng);
        if (width == -1) width = randRange(5, 9, rng);
        height = 5;
    }
    void generate() override { addFloor(); addWalls(); }
};

template <typename T> T triangularNumber(T n) { return n * (n + 1) / 2; }  // util/math_utils.hpp

class RearrangePlatform : public EmptyPlatform {  // scenario_rearrange.cpp:11-34
public:
    RearrangePlatform(Node *parent, Rng &rng, int walls, const FloatParams &params) : EmptyPlatform(parent, rng, walls, params) {}
    void init() override { height = randRange(4, 7, rng); length = 19; width = 14; }
    void generate() override { EmptyPlatform::generate(); }
};

class WallPlatform : public EmptyPlatform {  // platforms.hpp:332-373
public:
    WallPlatform(Node *parent, Rng &rng, int walls, const FloatParams &params, int w = -1) : EmptyPlatform(parent, rng, walls, params, w) {}
    void init() override {
        EmptyPlatform::init();
        wallHeight = randRange(param("obstaclesMinHeight"), param("obstaclesMaxHeight") + 1, rng);
        height = randRange(wallHeight + 4, wallHeight + 6, rng);
    }
    void generate() override {
        EmptyPlatform::generate();
        const auto wallX = randRange(1, length, rng);
        const auto wallThickness = randRange(1, length - wallX + 1, rng);
        layoutBoxes.emplace_back(makeAABB({wallX, 1, 1, wallX + wallThickness, 1 + wallHeight, width - 1}));
        for (int x = wallX; x < wallX + wallThickness; ++x)
            for (int z = 1; z < width; ++z) occupancy[{x, z}] = wallHeight;
    }
    int requiresMovableBoxesToTraverse() override { return triangularNumber(wallHeight - 1); }
    bool isMaxDifficulty() const override { return wallHeight >= param("obstaclesMaxHeight"); }
    int wallHeight{};
};

class LavaPlatform : public EmptyPlatform {  // platforms.hpp:375-407
public:
    LavaPlatform(Node *parent, Rng &rng, int walls, const FloatParams &params, int w = -1) : EmptyPlatform(parent, rng, walls, params, w) {}
    void init() override {
        EmptyPlatform::init();
        length = randRange(6, 12, rng);
        auto minLava = std::min(param("obstaclesMinLava"), length - 2);
        auto maxLava = std::min(param("obstaclesMaxLava") + 1, length - 1);
        lavaLength = randRange(minLava, maxLava, rng);
    }
    void generate() override {
        EmptyPlatform::generate();
        const auto lavaX = randRange(1, length - lavaLength, rng);
        terrainBoxes[TERRAIN_LAVA].emplace_back(makeAABB({lavaX, 1, 1, lavaX + lavaLength, 2, width - 1}));
    }
    int requiresMovableBoxesToTraverse() override { return std::max(1, lavaLength - 1); }
    bool isMaxDifficulty() const override { return lavaLength >= param("obstaclesMaxLava"); }
    int lavaLength{};
};

class StepPlatform : public EmptyPlatform {  // platforms.hpp:409-456
public:
    StepPlatform(Node *parent, Rng &rng, int walls, const FloatParams &params, int w = -1) : EmptyPlatform(parent, rng, walls, params, w) {}
    void init() override {
        EmptyPlatform::init();
        stepHeight = randRange(param("obstaclesMinHeight"), param("obstaclesMaxHeight") + 1, rng);
        height = randRange(stepHeight + 2, stepHeight + 5, rng);
    }
    void generate() override {
        const auto stepX = randRange(1, length, rng);
        layoutBoxes.emplace_back(makeAABB({0, 0, 0, stepX + 1, 1, width}));
        layoutBoxes.emplace_back(makeAABB({stepX, stepHeight, 0, length, stepHeight + 1, width}));
        layoutBoxes.emplace_back(makeAABB({stepX, 0, 0, stepX + 1, stepHeight + 1, width}));
        nextPlatformAnchor = newNode(root);
        nextPlatformAnchor->translateLocal({float(length), float(stepHeight), 0});
        addWalls();
        for (int x = stepX + 1; x < length; ++x)
            for (int z = 1; z < width; ++z) occupancy[{x, z}] = stepHeight;
    }
    int requiresMovableBoxesToTraverse() override { return triangularNumber(stepHeight - 1); }
    bool isMaxDifficulty() const override { return stepHeight >= param("obstaclesMaxHeight"); }
    int stepHeight{};
};

class GapPlatform : public EmptyPlatform {  // platforms.hpp:459-514
public:
    GapPlatform(Node *parent, Rng &rng, int walls, const FloatParams &params, int w = -1) : EmptyPlatform(parent, rng, walls, params, w) {}
    void init() override {
        EmptyPlatform::init();
        gap = randRange(param("obstaclesMinGap"), std::min(param("obstaclesMaxGap") + 1, length - 1), rng);
        gapX = randRange(1, length - gap, rng);
    }
    void generate() override {
        layoutBoxes.emplace_back(makeAABB({0, 0, 0, gapX, 1, width}));
        layoutBoxes.emplace_back(makeAABB({gapX + gap, 0, 0, length, 1, width}));
        nextPlatformAnchor = newNode(root);
        nextPlatformAnchor->translateLocal({float(length), 0, 0});
        addWalls();
    }
    int requiresMovableBoxesToTraverse() override { return triangularNumber(std::max(0, gap - 2)); }
    std::vector<VoxelCoords> generateObjectPositions(int n) override {
        std::vector<VoxelCoords> boxes, candidates;
        for (int x = 0; x < length; ++x)
            for (int z = 1; z < width - 1; ++z) {
                if (x >= gapX && x < gapX + gap) continue;
                candidates.emplace_back(x, 1, z);
            }
        for (int i = 0; i < n; ++i) {
            const auto v = randomSample(candidates, rng);
            const int y = ++occupancy[{v.x, v.z}];
            boxes.emplace_back(v.x, y, v.z);
        }
        return adjustTransformation(boxes);
    }
    int gap{}, gapX{};
};

class StartPlatform : public EmptyPlatform {  // platforms.hpp:516-528
public:
    StartPlatform(Node *parent, Rng &rng, const FloatParams &params, int w = -1) : EmptyPlatform(parent, rng, WALLS_SOUTH | WALLS_EAST | WALLS_WEST, params, w) {}
};

class ExitPlatform : public EmptyPlatform {  // platforms.hpp:530-546
public:
    ExitPlatform(Node *parent, Rng &rng, const FloatParams &params, int w = -1) : EmptyPlatform(parent, rng, WALLS_NORTH | WALLS_EAST | WALLS_WEST, params, w) {}
    void generate() override {
        EmptyPlatform::generate();
        terrainBoxes[TERRAIN_EXIT].emplace_back(makeAABB({length - 3, 1, 1, length - 1, 3, width - 1}));
    }
};

class TransitionPlatform : public EmptyPlatform {  // platforms.hpp:548-558
public:
    TransitionPlatform(Node *parent, Rng &rng, int walls, const FloatParams &params, int l, int w) : EmptyPlatform(parent, rng, walls, params, -1) { length = l; width = w; }
    void init() override { height = 5; }
};

// siv::PerlinNoise (src/libs/util/include/util/perlin_noise.hpp:56-322, Ken Perlin's improved noise, double precision):
// permutation = std::shuffle of 0..255 with std::default_random_engine(seed) (:118-131)
class PerlinNoise {
public:
    explicit PerlinNoise(std::uint32_t seed) {
        for (size_t i = 0; i < 256; ++i) p[i] = static_cast<std::uint8_t>(i);
        std::shuffle(std::begin(p), std::begin(p) + 256, std::default_random_engine(seed));
        for (size_t i = 0; i < 256; ++i) p[256 + i] = p[i];
    }
    static double Fade(double t) { return t * t * t * (t * (t * 6 - 15) + 10); }
    static double Lerp(double t, double a, double b) { return a + t * (b - a); }
    static double Grad(std::uint8_t hash, double x, double y, double z) {
        const std::uint8_t h = hash & 15;
        const double u = h < 8 ? x : y;
        const double v = h < 4 ? y : h == 12 || h == 14 ? x : z;
        return ((h & 1) == 0 ? u : -u) + ((h & 2) == 0 ? v : -v);
    }
    double noise3D(double x, double y, double z) const {
        const std::int32_t X = static_cast<std::int32_t>(std::floor(x)) & 255;
        const std::int32_t Y = static_cast<std::int32_t>(std::floor(y)) & 255;
        const std::int32_t Z = static_cast<std::int32_t>(std::floor(z)) & 255;
        x -= std::floor(x); y -= std::floor(y); z -= std::floor(z);
        const double u = Fade(x), v = Fade(y), w = Fade(z);
        const std::int32_t A = p[X] + Y, AA = p[A] + Z, AB = p[A + 1] + Z;
        const std::int32_t B = p[X + 1] + Y, BA = p[B] + Z, BB = p[B + 1] + Z;
        return Lerp(w, Lerp(v, Lerp(u, Grad(p[AA], x, y, z), Grad(p[BA], x - 1, y, z)), Lerp(u, Grad(p[AB], x, y - 1, z), Grad(p[BB], x - 1, y - 1, z))),
                    Lerp(v, Lerp(u, Grad(p[AA + 1], x, y, z - 1), Grad(p[BA + 1], x - 1, y, z - 1)),
                         Lerp(u, Grad(p[AB + 1], x, y - 1, z - 1), Grad(p[BB + 1], x - 1, y - 1, z - 1))));
    }
    double accumulatedOctaveNoise2D_0_1(double x, double y, std::int32_t octaves) const {
        double result = 0, amp = 1;
        for (std::int32_t i = 0; i < octaves; ++i) {
            result += noise3D(x, y, 0) * amp;
            x *= 2; y *= 2; amp /= 2;
        }
        return std::clamp<double>(result * 0.5 + 0.5, 0, 1);
    }
    std::uint8_t p[512];
};

struct BBoxInfo {  // component_voxel_grid.hpp:33-50
    uint8_t type{};
    ColorRgb color{};
    bool operator<(const BBoxInfo &o) const { return type == o.type ? color < o.color : type < o.type; }
};
using Boxes = std::vector<BoundingBox>;

struct CoordRange { int min, max; };
inline CoordRange startEndCoord(int bmin, int bmax, int dir) {
    if (dir == 1) return {bmax + 1, bmax + 1};
    else if (dir == -1) return {bmin - 1, bmin - 1};
    else return {bmin, bmax};
}

class VoxelGridComponent {  // component_voxel_grid.hpp:61-190
public:
    explicit VoxelGridComponent(int maxVoxelsXYZ = 100, float minX = 0, float minY = 0, float minZ = 0, float voxelSize = 1)
        : grid(size_t(maxVoxelsXYZ), Vec3{minX, minY, minZ}, voxelSize) {}
    void reset() { grid.clear(); }
    static Voxel makeVoxel(int type, int terrain = 0, ColorRgb color = LAYOUT_DEFAULT) { Voxel v; v.voxelType = uint8_t(type); v.terrain = uint8_t(terrain); v.color = color; return v; }
    void addPlatform(const Platform &p, ColorRgb layoutColor, ColorRgb wallColor, bool drawWalls = true) {
        for (auto &bb : p.layoutBoxes) addBoundingBox(bb.boundingBox(), makeVoxel(1 | 2, TERRAIN_NONE, layoutColor));
        for (auto &bb : p.wallBoxes) addBoundingBox(bb.boundingBox(), makeVoxel(1 | (int(drawWalls) << 1), TERRAIN_NONE, wallColor));
        for (auto &[terrainType, v] : p.terrainBoxes)
            for (auto &bb : v) addTerrainBoundingBox(bb.boundingBox(), terrainType);
    }
    void addBoundingBox(const BoundingBox &bb, const Voxel &v) {
        for (int x = bb.min.x; x < bb.max.x; ++x)
            for (int y = bb.min.y; y < bb.max.y; ++y)
                for (int z = bb.min.z; z < bb.max.z; ++z) grid.set({x, y, z}, v);
    }
    void addTerrainBoundingBox(const BoundingBox &bb, int terrain) {
        for (int x = bb.min.x; x < bb.max.x; ++x)
            for (int y = bb.min.y; y < bb.max.y; ++y)
                for (int z = bb.min.z; z < bb.max.z; ++z) {
                    const VoxelCoords c{x, y, z};
                    if (!grid.hasVoxel(c)) grid.set(c, Voxel());
                    grid.get(c)->terrain |= terrain;
                }
    }
    std::map<BBoxInfo, Boxes> toBoundingBoxes() {
        static const int directions[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        std::unordered_set<VoxelCoords, VoxelHash> visited;
        const auto gridHashMap = grid.getHashMap();  // a COPY, as in the reference (:114)
        std::map<BBoxInfo, Boxes> boxesByVoxelType;
        for (auto it : gridHashMap) {
            const auto &coord = it.first;
            const auto &voxel = it.second;
            const auto voxelType = voxel.voxelType;
            const auto color = voxel.color;
            if (visited.count(coord)) continue;
            visited.emplace(coord);
            BoundingBox bbox{coord, coord};
            std::vector<VoxelCoords> expansion;
            for (auto &direction : directions)
                for (int sign = -1; sign <= 1; sign += 2) {
                    const int dx = direction[0] * sign, dy = direction[1] * sign, dz = direction[2] * sign;
                    bool canExpand = true;
                    while (true) {
                        const auto xlim = startEndCoord(bbox.min.x, bbox.max.x, dx);
                        const auto ylim = startEndCoord(bbox.min.y, bbox.max.y, dy);
                        const auto zlim = startEndCoord(bbox.min.z, bbox.max.z, dz);
                        expansion.clear();
                        for (auto x = xlim.min; x <= xlim.max && canExpand; ++x)
                            for (auto y = ylim.min; y <= ylim.max && canExpand; ++y)
                                for (auto z = zlim.min; z <= zlim.max; ++z) {
                                    const VoxelCoords c{x, y, z};
                                    const auto v = grid.get(c);
                                    if (!v || v->voxelType != voxelType || v->color != color || visited.count(c)) { canExpand = false; break; }
                                    expansion.emplace_back(c);
                                }
                        if (!canExpand) break;
                        for (auto nc : expansion) { visited.emplace(nc); bbox.addPoint(nc); }
                    }
                }
            boxesByVoxelType[{voxelType, color}].emplace_back(bbox);
        }
        return boxesByVoxelType;
    }
    VoxelGrid grid;
};

}  // namespace orc

namespace std {
template <> struct hash<orc::VoxelCoords> {
    size_t operator()(const orc::VoxelCoords &v) const noexcept { return orc::VoxelHash{}(v); }
};
}  // namespace std
