// ORACLE (test infrastructure, NOT product code): one environment = state + agents + scenario logic.
// Restated from
//   src/libs/env/src/env.cpp:36-152 (ctor, reset, step)      src/libs/env/include/env/env.hpp:124-170 (EnvState)
//   src/libs/env/src/agent.cpp:26-170 (DefaultKinematicAgent)  src/libs/env/include/env/physics.hpp:58-85 (RigidBody)
//   src/libs/env/include/env/scenario.hpp:94-117,184-307 (reward shaping, rewardAgent/Team/All, doneWithTimer)
//   src/libs/scenarios/include/scenarios/scenario_default.hpp:80-186 (spawnAgents, agent drawables, HUD)
//   src/libs/scenarios/include/scenarios/component_object_stacking.hpp:45-198
//   src/libs/scenarios/include/scenarios/component_fall_detection.hpp:33-56
//   src/libs/scenarios/src/layout_utils.cpp:17-68 (addBoundingBoxes, addTerrain)
//   src/libs/scenarios/src/scenario_tower_building.cpp:8-266 (+ scenario_tower_building.hpp:42-52)
//   src/libs/env/src/vector_env.cpp:89-120 (done handling)   src/libs/bindings/megaverse.cpp:60-69,100-137
//   + the other scenario sources (obstacles, collect, sokoban, rearrange, hex_explore, hex_memory, component_hexagonal_maze), cited inline.
// PINNED: the reference's whole env library (env.cpp, agent.cpp, kinematic_character_controller.cpp, vector_env.cpp, every scenario
// source) is compiled in place on a Bullet stand-in (oracle/ref_shim/env_shim.cpp, mini_bullet/) and this restatement follows it bit
// for bit, tick by tick (tests/test_ref_shim.py, tests/test_ref_golden.py).  What that leaves unpinned is stated in orc_physics.hpp.
#pragma once
#include <cfloat>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <queue>
#include <set>
#include <stdexcept>

#include "orc_level.hpp"
#include "orc_maze.hpp"
#include "orc_physics.hpp"

namespace orc {

enum Action {
    A_Idle = 0, A_Left = 1 << 1, A_Right = 1 << 2, A_Forward = 1 << 3, A_Backward = 1 << 4, A_LookLeft = 1 << 5, A_LookRight = 1 << 6,
    A_Jump = 1 << 7, A_Interact = 1 << 8, A_LookDown = 1 << 9, A_LookUp = 1 << 10,
};
enum MeshType { MESH_BOX = 0, MESH_CAPSULE = 1, MESH_SPHERE = 2, MESH_CONE = 3, MESH_CYLINDER = 4, MESH_NUM = 5 };

struct Instance {  // one drawable: mesh type, palette colour, model matrix (absoluteTransformationMatrix)
    int mesh;
    int color;  // palette index
    Mat4 model;
};

// DrawablesMap (env.hpp:57-67,72): mesh type -> drawables in insertion order.  An entry names the scene-graph object whose
// absolute transformation is the model matrix at draw time.
struct DrawEntry {
    enum Kind { D_STATIC, D_OBJECT, D_EYES, D_BAR, D_BODY, D_REWARD_ROOT, D_REWARD_BOTTOM, D_MEMORY } kind;  // D_MEMORY: index = object * 4 + child
    int index;
    int color;
    Mat4 model;  // D_STATIC
};

struct Agent {
    KCC kcc;
    float currXRotation = 0.0f;
    float verticalLookLimitRad = 0.2f;
    Mat4 objectT = mat4Identity();      // the agent Object3D transformation
    Mat4 cameraLocal = mat4Identity();  // cameraObject, child of agent
    Mat4 pickupLocal = mat4Identity();  // pickupSpot, child of cameraObject
    Mat4 bodyLocal, eyesLocal, uiLocal, barAnchorLocal, barLocal;
    int color = 0;
    Vec3 spawnPos{0, 0, 0};  // constructor arguments, kept for the scenario pin (tests/test_ref_shim.py)
    float spawnRot = 0.0f;

    static constexpr float rotateRadians = 3.5f, rotateXRadians = 1.5f, agentHeight = 1.75f;

    void init(Vec3 startingPosition, float rotationRad, float lookLimit) {  // agent.cpp:26-67
        *this = Agent{};
        spawnPos = startingPosition, spawnRot = rotationRad;
        verticalLookLimitRad = lookLimit;
        cameraLocal = mul(mat4Translation({0, 0.41f, 0}), cameraLocal);
        pickupLocal = mul(mat4Translation({0.0f, -0.44f, -1.0f}), pickupLocal);
        kcc.basis = mat3FromQuat(quatAxisAngle({0, 1, 0}, rotationRad));
        // KinematicCharacterController's constructor then runs setUp(up) -> setGravity -> setUpVector (kinematic_character_controller.cpp:
        // 148,646-651,727-749): with the up axis unchanged the correcting rotation is the identity quaternion, but the ghost's basis
        // still goes matrix -> quaternion -> matrix once (xform.getRotation(), xform.setRotation(orn)), which is not always a no-op in float
        kcc.basis = mat3FromQuat(quatFromMat3(kcc.basis));
        kcc.pos = {startingPosition.x, startingPosition.y + agentHeight, startingPosition.z};
    }
    void updateTransform() {  // agent.cpp:73-98
        const Quat q = quatFromMat3(kcc.basis);
        Vec3 position = kcc.pos;
        const Vec3 axis = quatAxis(q);
        const Vec3 normalizedAxis = mgNormalized(axis);
        const float rotation = quatAngle(q);
        if (std::isnan(position.x) || std::isnan(position.y) || std::isnan(position.z) || std::isnan(rotation)) return;
        if (std::isnan(normalizedAxis.x) || std::isnan(normalizedAxis.y) || std::isnan(normalizedAxis.z)) return;
        position += Vec3{0, 0.05f, 0.0f};
        objectT = mul(mat4Translation(position), mul(mat4Rotation(rotation, normalizedAxis), mat4Identity()));
    }
    void rotateYAxis(float radians) { kcc.basis = mul(kcc.basis, mat3FromQuat(quatAxisAngle({0, 1, 0}, radians))); }  // agent.cpp:128-133
    void lookLeft(float dt) { rotateYAxis(rotateRadians * dt); }
    void lookRight(float dt) { rotateYAxis(-rotateRadians * dt); }
    void lookUp(float dt) {  // agent.cpp:110-116
        cameraLocal = mul(cameraLocal, mat4RotationX(-currXRotation));
        currXRotation += rotateXRadians * dt;
        currXRotation = std::min(verticalLookLimitRad, currXRotation);
        cameraLocal = mul(cameraLocal, mat4RotationX(currXRotation));
    }
    void lookDown(float dt) {  // agent.cpp:118-126
        cameraLocal = mul(cameraLocal, mat4RotationX(-currXRotation));
        currXRotation -= rotateXRadians * dt * 1.1f;
        currXRotation = std::max(-verticalLookLimitRad, currXRotation);
        cameraLocal = mul(cameraLocal, mat4RotationX(currXRotation));
    }
    Vec3 forwardDirection() const { Vec3 f = kcc.basis.r[2]; f.z = -f.z; return btNormalized(f); }      // agent.cpp:135-142
    Vec3 strafeLeftDirection() const { Vec3 s = kcc.basis.r[0]; s.x = -s.x; return btNormalized(s); }  // agent.cpp:144-150
    bool onGround() const { return kcc.onGround(); }
    void jump() { if (onGround()) kcc.jump({0, 6.2f, 0}); }
    Mat4 cameraAbs() const { return mul(objectT, cameraLocal); }                         // Object::absoluteTransformation, left to right
    Mat4 pickupAbs() const { return mul(mul(objectT, cameraLocal), pickupLocal); }
};

struct MovableObject {
    Mat4 local;            // transformation relative to parent
    int parentAgent = -1;  // -1: scene; else child of that agent's pickupSpot
    Vec3 collisionScale{1.15f, 1.15f, 1.15f}, collisionOffset{0, -0.05f, 0};
    int collider = -1;
    int color = 0;
    int mesh = MESH_BOX;
    ColorRgb arrangementColor = WHITE;
    bool pickedUp = false;  // ArrangementObject (scenario_rearrange.hpp:60-71)
};

struct StaticBox { BoundingBox bb; uint8_t type; ColorRgb color; };
struct TerrainSlab { int terrain; BoundingBox bb; };

using RewardShaping = std::map<std::string, float>;

class Env {
public:
    enum Scenario { S_TOWER = 0, S_OBSTACLES = 1, S_COLLECT = 2, S_REARRANGE = 3, S_SOKOBAN = 4, S_HEX_EXPLORE = 5, S_HEX_MEMORY = 6, S_EMPTY = 7 };
    enum SokobanTerrain { SOKO_EMPTY = 0, SOKO_WALL = 1, SOKO_GOAL = 2 };
    struct SokobanLevel { std::vector<std::string> rows; };
    enum PlatformType { PT_EMPTY, PT_WALL, PT_LAVA, PT_STEP, PT_GAP };
    struct RewardObject { Mat4 root, bottomLocal; int color; };
    struct MemoryObject { Mat4 root; Mat4 childLocal[2]; int numChildren = 0; bool good = false, alive = true; VoxelCoords voxel{0, 0, 0}; };  // CollectableObject
    struct ArrangementItem { int shape = MESH_SPHERE; ColorRgb color = WHITE; VoxelCoords offset{0, 0, 0}; };  // scenario_rearrange.hpp:20-53

    Env(const std::string &scenarioName, int numAgents, const FloatParams &custom) : numAgents(numAgents) {
        std::string n;
        for (char ch : scenarioName) n.push_back(char(tolower(ch)));
        // Scenario::init (scenario.hpp:98-107) + initializeDefaultParameters (:225-231)
        floatParams["episodeLengthSec"] = 60.0f;
        floatParams["verticalLookLimitRad"] = 0.2f;
        floatParams["useUIRewardIndicators"] = 0.0f;
        if (n == "towerbuilding") scenario = S_TOWER;
        else if (n == "collect") scenario = S_COLLECT;
        else if (n == "rearrange") scenario = S_REARRANGE;
        else if (n == "hexexplore") scenario = S_HEX_EXPLORE;
        else if (n == "hexmemory") scenario = S_HEX_MEMORY;
        else if (n == "empty") scenario = S_EMPTY;  // scenario_empty.cpp
        else if (n == "sokoban") {  // scenario_sokoban.cpp:39-81, scenario_sokoban.hpp:50-54
            scenario = S_SOKOBAN;
            floatParams["episodeLengthSec"] = 80.0f;
            vg = VoxelGridComponent(100, 0, 0, 0, 2);
            const char *envvar = std::getenv("BOXOBAN_LEVELS");
            std::string dir = (envvar && std::strlen(envvar)) ? envvar : "~/datasets/boxoban";
            const auto tilde = dir.find('~');
            if (tilde != std::string::npos) {
                const char *home = std::getenv("HOME");
                if (!home || !std::strlen(home)) throw std::runtime_error("oracle: HOME not set");
                dir.replace(tilde, 1, home);
            }
            const std::string dirWithLevels = dir + "/unfiltered/train";
            for (int levelFileIdx = 0; levelFileIdx <= 999; ++levelFileIdx) {
                char name[16];
                std::snprintf(name, sizeof name, "%03d.txt", levelFileIdx);
                const std::string path = dirWithLevels + "/" + name;
                if (std::ifstream(path).good()) allSokobanLevelFiles.emplace_back(path);
            }
            if (allSokobanLevelFiles.empty()) throw std::runtime_error("oracle: no Boxoban levels (set BOXOBAN_LEVELS)");
        }
        else if (n.rfind("obstacles", 0) == 0 || n == "test") {
            // ObstaclesScenario::initializeDefaultParameters + the registered variants (scenario_obstacles.hpp:48-270, init.hpp:33-56)
            scenario = S_OBSTACLES;
            platformTypes = {PT_WALL, PT_LAVA, PT_STEP, PT_GAP};
            auto &fp = floatParams;
            fp["obstaclesMinNumPlatforms"] = 1; fp["obstaclesMaxNumPlatforms"] = 2; fp["obstaclesMinGap"] = 1; fp["obstaclesMaxGap"] = 2;
            fp["obstaclesMinLava"] = 1; fp["obstaclesMaxLava"] = 4; fp["obstaclesMinHeight"] = 1; fp["obstaclesMaxHeight"] = 3;
            fp["obstaclesNumAllowedMaxDifficulty"] = 1;
            if (n == "obstacleseasy" || n == "obstacles") {
            } else if (n == "obstaclesmedium") {
                fp["obstaclesMinNumPlatforms"] = 2; fp["obstaclesMaxNumPlatforms"] = 4; fp["obstaclesMinLava"] = 2; fp["obstaclesMaxLava"] = 5;
            } else if (n == "obstacleshard") {
                fp["obstaclesMinNumPlatforms"] = 2; fp["obstaclesMaxNumPlatforms"] = 7; fp["obstaclesMinGap"] = 2; fp["obstaclesMaxGap"] = 3;
                fp["obstaclesMinLava"] = 3; fp["obstaclesMaxLava"] = 10; fp["obstaclesMinHeight"] = 2; fp["obstaclesMaxHeight"] = 4;
            } else if (n == "obstacleswalls" || n == "obstaclessteps" || n == "obstacleslava") {
                fp["obstaclesMinNumPlatforms"] = 1; fp["obstaclesMaxNumPlatforms"] = 4; fp["obstaclesMinGap"] = 1; fp["obstaclesMaxGap"] = 3;
                fp["obstaclesMinLava"] = 2; fp["obstaclesMaxLava"] = 10; fp["obstaclesMinHeight"] = 1; fp["obstaclesMaxHeight"] = 3;
                platformTypes = {n == "obstacleswalls" ? PT_WALL : (n == "obstaclessteps" ? PT_STEP : PT_LAVA)};
                onePlatformType = true;
            } else if (n == "test") {
                fp["obstaclesMinNumPlatforms"] = 0; fp["obstaclesMaxNumPlatforms"] = 0; fp["episodeLengthSec"] = 6.0f;
            } else throw std::runtime_error("oracle: unknown scenario " + n);
        } else throw std::runtime_error("oracle: unknown scenario " + n);
        rewardShaping.assign(size_t(numAgents), RewardShaping{{"teamSpirit", 0.0f}});
        for (auto &rs : rewardShaping)
            for (auto &[k, v] : defaultRewardShaping()) rs[k] = v;
        for (auto &[k, v] : custom) floatParams[k] = v;
        currAction.assign(size_t(numAgents), 0);
        lastReward.assign(size_t(numAgents), 0.0f);
        totalReward.assign(size_t(numAgents), 0.0f);
        agentState.assign(size_t(numAgents), TowerAgentState{});
    }

    RewardShaping defaultRewardShaping() const {
        if (scenario == S_COLLECT)  // scenario_collect.hpp:42-50
            return {{"collectSingleGood", 1.0f}, {"collectSingleBad", -1.0f}, {"collectAll", 5.0f}, {"collectAbyss", -0.5f}};
        if (scenario == S_OBSTACLES)  // scenario_obstacles.hpp:37-45,201-206
            return {{"obstaclesAgentAtExit", 1.0f}, {"obstaclesAllAgentsAtExit", 5.0f}, {"obstaclesExtraReward", 0.5f},
                    {"obstaclesAgentCarriedObjectToExit", onePlatformType ? 1.0f : 0.0f}};
        if (scenario == S_HEX_EXPLORE) return {{"exploreSolved", 5.0f}};  // scenario_hex_explore.hpp:27-30
        if (scenario == S_HEX_MEMORY) return {{"memoryCollectGood", 1.0f}, {"memoryCollectBad", -1.0f}};  // scenario_hex_memory.hpp:43-49
        if (scenario == S_SOKOBAN)  // scenario_sokoban.hpp:41-48
            return {{"sokobanBoxOnTarget", 1.0f}, {"sokobanBoxLeavesTarget", -1.0f}, {"sokobanAllBoxesOnTarget", 10.0f}};
        if (scenario == S_EMPTY) return {};  // scenario_empty.hpp:28
        if (scenario == S_REARRANGE)  // scenario_rearrange.hpp:91-97
            return {{"rearrangeOneMoreObjectCorrectPosition", 1.0f}, {"rearrangeAllObjectsCorrectPosition", 10.0f}};
        // scenario_tower_building.hpp:44-52
        return {{"teamSpirit", 0.1f}, {"towerPickedUpObject", 0.1f}, {"towerVisitedBuildingZoneWithObject", 0.1f}, {"towerBuildingReward", 1.0f}};
    }

    void seed(int s) { rng.seed((unsigned long)s); }

    // ---------------------------------------------------------------- reset (env.cpp:57-76)
    void reset() {
        done = false; currEpisodeSec = 0; numFrames = 0;
        std::fill(currAction.begin(), currAction.end(), 0);
        std::fill(lastReward.begin(), lastReward.end(), 0.0f);
        std::fill(totalReward.begin(), totalReward.end(), 0.0f);
        agents.clear(); colliders.clear(); objects.clear(); staticBoxes.clear(); terrainSlabs.clear(); rewardObjects.clear(); drawables.clear();

        auto sd = randRange(0, 1 << 30, rng);
        rng.seed((unsigned long)sd);
        episodeSeed = unsigned(sd);

        if (scenario == S_TOWER) towerReset(); else if (scenario == S_OBSTACLES) obstaclesReset(); else if (scenario == S_COLLECT) collectReset();
        else if (scenario == S_REARRANGE) rearrangeReset(); else if (scenario == S_SOKOBAN) sokobanReset(); else if (scenario == S_HEX_EXPLORE) hexExploreReset();
        else if (scenario == S_EMPTY) {  // EmptyScenario::reset / agentStartingPositions (scenario_empty.cpp:15-22); no components
            agentSpawnPositions.assign(size_t(numAgents), Vec3{1, 1, 1});
            carryingObject.assign(size_t(numAgents), -1);  // (only read by the state dump)
        }
        else hexMemoryReset();
        if (scenario == S_HEX_MEMORY) hexMemorySpawnAgents(); else
        spawnAgents();
        if (scenario == S_TOWER) towerAddEpisodeDrawables(); else if (scenario == S_OBSTACLES) obstaclesAddEpisodeDrawables();
        else if (scenario == S_COLLECT) collectAddEpisodeDrawables(); else if (scenario == S_REARRANGE) rearrangeAddEpisodeDrawables();
        else if (scenario == S_SOKOBAN) sokobanAddEpisodeDrawables(); else if (scenario == S_HEX_EXPLORE) hexExploreAddEpisodeDrawables();
        else if (scenario == S_EMPTY) addStaticCollidingBox(Vec3{10, 1, 10}, Vec3{5, 0, 5}, BLUE);  // scenario_empty.cpp:25-28
        else hexMemoryAddEpisodeDrawables();
        addAgentsAndUI();
    }

    // ---------------------------------------------------------------- hexagonal maze component (component_hexagonal_maze.cpp:19-128)
    void hexMazeReset(int minSize, int maxSize, float omitMin, float omitMax) {
        hexMazeSize = randRange(minSize, maxSize, rng);
        hexMaze = std::make_unique<HoneyCombMaze>(hexMazeSize);
        hexMaze->initialiseGraph();
        std::mt19937 generator(episodeSeed ^ 0x6d617a65u);  // upstream: std::random_device (see orc_maze.hpp)
        hexMaze->generate(generator);
        const auto b = hexMaze->coordinateBounds();
        hexXMin = b[0], hexYMin = b[1], hexXMax = b[2], hexYMax = b[3];
        hexMazeScale = 3.5f;
        hexWallHeight = frand(rng) * 0.55f + 0.85f;
        hexOmitWallsProbability = frand(rng) * (omitMax - omitMin) + omitMin;
        hexWallLandmarkProbability = frand(rng) * 0.15f + 0.15f;
        hexBottomEdgingColor = sampleRandomColor(rng);
        hexTopEdgingColor = sampleRandomColor(rng);
        hexXMin *= hexMazeScale, hexXMax *= hexMazeScale, hexYMin *= hexMazeScale, hexYMax *= hexMazeScale;
    }
    void hexMazeAddDrawablesAndCollisions() {
        const Vec3 scale{float(hexXMax - hexXMin), 0.0001f, float(hexYMax - hexYMin)};
        const Vec3 translation{float(hexXMax + hexXMin) / 2, 0.0f, float(hexYMax + hexYMin) / 2};
        addStaticCollidingBox(scale, translation, randomLayoutColor(rng));
        std::set<std::pair<int, int>> existingWalls;
        const auto &adjList = hexMaze->adjacency;
        for (int cellIdx = 0; cellIdx < int(adjList.size()); ++cellIdx) {
            for (const auto &entry : adjList[size_t(cellIdx)]) {
                const int adjCellIdx = entry.cell;
                auto cellPair = std::make_pair(cellIdx, adjCellIdx);
                if (cellPair.first > cellPair.second) std::swap(cellPair.first, cellPair.second);
                if (adjCellIdx != -1) {
                    if (existingWalls.count(cellPair)) continue;
                    if (frand(rng) < hexOmitWallsProbability) continue;
                }
                existingWalls.insert(cellPair);
                double x1 = entry.border[0], z1 = entry.border[1], x2 = entry.border[2], z2 = entry.border[3];
                x1 *= hexMazeScale, z1 *= hexMazeScale, x2 *= hexMazeScale, z2 *= hexMazeScale;
                const auto length = 0.5f * sqrtf(float((x1 - x2) * (x1 - x2) + (z1 - z2) * (z1 - z2)));
                const Vec3 wallScale{length, hexWallHeight, 0.15f};
                const Vec3 wallTranslation{float(x1 + x2) / 2, hexWallHeight, float(z1 + z2) / 2};
                const auto deltaX = x1 - x2;
                const auto deltaZ = z1 - z2;
                float rotationY = float(M_PI_2);
                if (std::fabs(deltaX) > 1e-5f) {
                    const auto tanAlpha = deltaZ / deltaX;
                    rotationY = -atanf(float(tanAlpha));
                }
                // the wall object gets its transformation after its landmark children are created; absolute matrices are what count
                const Mat4 wallLocal = mul(mat4Translation(wallTranslation), mul(mat4RotationY(rotationY), mul(mat4Scaling(wallScale), mat4Identity())));
                if (frand(rng) < hexWallLandmarkProbability) {
                    const auto landmarkWidth = 0.15f, landmarkHeight = landmarkWidth * length / hexWallHeight;
                    int numLandmarks = randRange(2, 5, rng);
                    for (int li = 0; li < numLandmarks; ++li) {
                        const Vec3 landmarkScale{landmarkWidth, landmarkHeight, frand(rng) * 1.2f + 1.5f};
                        const Vec3 landmarkTranslation{float(li % 2 == 1) * landmarkWidth * 2, float(li > 1) * landmarkHeight * 2 - 0.2f, 0};
                        const Mat4 landmarkLocal = mul(mat4Translation(landmarkTranslation), mul(mat4Identity(), mat4Scaling(landmarkScale)));  // scaleLocal, translate
                        const int color = paletteIndex(sampleRandomColor(rng));
                        drawables[MESH_BOX].push_back({DrawEntry::D_STATIC, 0, color, mul(wallLocal, landmarkLocal)});
                    }
                }
                drawables[MESH_BOX].push_back({DrawEntry::D_STATIC, 0, paletteIndex(DARK_BLUE), wallLocal});
                {   // RigidBody child with identity local: world transform = rotation part + translation, local scaling = column lengths
                    Collider c;
                    c.kind = 0;
                    c.c = translationOf(wallLocal) + Vec3{0, 0, 0};
                    c.h = scalingOf(wallLocal) * Vec3{1, 1, 1};
                    c.rotated = true;
                    const float lxInv = 1.0f / c.h.x;  // Matrix4::rotation(): column * lengthInverted() (Math/Vector.h:547,561)
                    c.ax = wallLocal.c[0][0] * lxInv; c.az = wallLocal.c[0][2] * lxInv;
                    colliders.push_back(c);
                }
                {
                    const Vec3 edgingScale{length * 1.02f, hexWallHeight * 0.12f, 0.2f};
                    const Vec3 bottomEdgingTranslation{wallTranslation.x, edgingScale.y, wallTranslation.z};
                    const Mat4 m = mul(mat4Translation(bottomEdgingTranslation), mul(mat4RotationY(rotationY), mul(mat4Scaling(edgingScale), mat4Identity())));
                    drawables[MESH_BOX].push_back({DrawEntry::D_STATIC, 0, paletteIndex(hexBottomEdgingColor), m});
                }
            }
        }
    }

    // ---------------------------------------------------------------- HexExplore (scenario_hex_explore.cpp:22-108)
    void hexExploreReset() {
        solved = false;
        vg.reset();
        carryingObject.assign(size_t(numAgents), -1);
        objectSpawnPositions.clear(); rewardSpawnPositions.clear();
        hexMazeReset(2, 8, 0.1f, 0.4f);
        const auto &centers = hexMaze->cellCenters;
        const auto randomCellIdx = randRange(0, int(hexMaze->adjacency.size()), rng);
        const auto cellCenter = centers[size_t(randomCellIdx)];
        rewardObjectCoords = Vec3{float(cellCenter.first) * hexMazeScale, 0, float(cellCenter.second) * hexMazeScale};
        // agentStartingPositions (:63-99), asked for by DefaultScenario::spawnAgents before its own draws
        agentSpawnPositions.clear();
        std::vector<int> cellIndices(hexMaze->adjacency.size(), 0);
        std::iota(cellIndices.begin(), cellIndices.end(), 0);
        std::shuffle(cellIndices.begin(), cellIndices.end(), rng);
        float furtherstDistance = 0;
        for (auto cellIdx : cellIndices) {
            const auto cc = centers[size_t(cellIdx)];
            const Vec3 spawnPos{float(cc.first) * hexMazeScale, float(0.1), float(cc.second) * hexMazeScale};
            const auto distance = length(rewardObjectCoords - spawnPos);
            const auto rotation = float(2 * M_PI / numAgents);
            if (distance > furtherstDistance) {
                agentSpawnPositions.clear();
                for (int i = 0; i < numAgents; ++i) {
                    const Vec3 delta{sinf(float(i) * rotation), 0, cosf(float(i) * rotation)};
                    agentSpawnPositions.emplace_back(spawnPos + delta);
                }
                furtherstDistance = distance;
            }
            if (distance > float(hexMazeSize) * hexMazeScale) break;
        }
        if (agentSpawnPositions.empty()) agentSpawnPositions.assign(size_t(numAgents), Vec3{0, 1, 0});
        agentInitialPositions = agentSpawnPositions;
    }
    void hexExploreAddEpisodeDrawables() {
        hexMazeAddDrawablesAndCollisions();
        const auto scale = 1.9f;
        RewardObject ro;
        ro.bottomLocal = mul(mat4Translation({0.0f, -1.0f, 0.0f}), mul(mat4Identity(), mat4RotationX(180.0f * 3.14159265358979323846f / 180.0f)));
        ro.root = mul(mat4Translation(rewardObjectCoords + Vec3{0, 1.2f, 0}), mul(mat4Scaling({0.17f * scale, 0.35f * scale, 0.17f * scale}), mat4Identity()));
        ro.color = paletteIndex(VIOLET);
        addRewardDrawables(ro);
        rewardObjects.push_back(ro);
        exploreRewardAlive = true;
    }
    void hexExploreStep() {
        for (int i = 0; i < numAgents; ++i) {
            const Vec3 t = translationOf(agents[size_t(i)].objectT);
            const auto threshold = 1.2;
            const auto distance = length(t - rewardObjectCoords);
            if (distance < threshold && !solved) {
                solved = true;
                doneWithTimer();
                rewardTeam("exploreSolved", i, 1);
                rewardObjects[0].root = mul(mat4Translation({1e3f, 1e3f, 1e3f}), rewardObjects[0].root);
                exploreRewardAlive = false;
                break;
            }
        }
    }

    // ---------------------------------------------------------------- HexMemory (scenario_hex_memory.cpp:19-217)
    enum MemoryShape { SHAPE_PILLAR, SHAPE_DIAMOND, SHAPE_SPHERE };
    void hexMemoryReset() {
        solved = false;
        vg.reset();
        carryingObject.assign(size_t(numAgents), -1);
        objectSpawnPositions.clear(); rewardSpawnPositions.clear();
        goodObjects.clear(), badObjects.clear();
        goodObjectsCollected = 0;
        memoryObjects.clear();
        hexMazeReset(2, 8, 0.1f, 0.95f);
        const auto &centers = hexMaze->cellCenters;
        const int numCells = int(hexMaze->adjacency.size());
        float minDistanceToCenter = 1e9f;
        int centerCellIdx = 0;
        for (int cellIdx = 0; cellIdx < numCells; ++cellIdx) {
            const auto cellCenter = centers[size_t(cellIdx)];
            const auto distanceToCenter = std::sqrt(cellCenter.first * cellCenter.first + cellCenter.second * cellCenter.second);
            if (distanceToCenter < minDistanceToCenter) { centerCellIdx = cellIdx; minDistanceToCenter = float(distanceToCenter); }
        }
        const auto mazeCenter = centers[size_t(centerCellIdx)];
        landmarkLocation = Vec3{float(mazeCenter.first * hexMazeScale), 1.0f, float(mazeCenter.second * hexMazeScale)};
        std::vector<Vec3> objectCoordinates;
        for (int cellIdx = 0; cellIdx < numCells; ++cellIdx) {
            if (cellIdx == centerCellIdx) continue;
            const auto cellCenter = centers[size_t(cellIdx)];
            Vec3 coord{float(cellCenter.first), 0.5f, float(cellCenter.second)};
            // Vector3(frand - 0.5f, 0, frand - 0.5f): GCC evaluates call arguments right to left, z's draw comes first
            const float oz = frand(rng) - 0.5f;
            const float ox = frand(rng) - 0.5f;
            coord += Vec3{ox, 0, oz};
            objectCoordinates.emplace_back(Vec3{coord.x * hexMazeScale, coord.y, coord.z * hexMazeScale});
        }
        std::shuffle(objectCoordinates.begin(), objectCoordinates.end(), rng);
        const float cellWithObjectsFraction = frand(rng) * 0.25f + 0.2f;
        const long numCellsWithGoodObjects = std::lround(ceilf(cellWithObjectsFraction * objectCoordinates.size()));
        const long numCellsWithBadObjects = std::lround(ceilf(cellWithObjectsFraction * objectCoordinates.size()));
        goodObjects = std::vector<Vec3>(objectCoordinates.begin(), objectCoordinates.begin() + numCellsWithGoodObjects);
        if (int(objectCoordinates.size()) >= numCellsWithGoodObjects + numCellsWithBadObjects)
            badObjects = std::vector<Vec3>(objectCoordinates.begin() + numCellsWithGoodObjects, objectCoordinates.begin() + numCellsWithGoodObjects + numCellsWithBadObjects);
        // agentStartingPositions (:121-130)
        const auto rotationBetweenAgents = float(2 * M_PI / numAgents);
        agentSpawnPositions.assign(size_t(numAgents), Vec3{0, 0, 0});
        for (int i = 0; i < numAgents; ++i)
            agentSpawnPositions[size_t(i)] = Vec3{sinf(rotationBetweenAgents * float(i)), float(0.3), cosf(rotationBetweenAgents * float(i))} * 1.5f;
        agentInitialPositions = agentSpawnPositions;
    }
    void hexMemorySpawnAgents() {  // :132-153: evenly spaced headings, no draws
        const float lookLimit = floatParams["verticalLookLimitRad"];
        const auto rotationBetweenAgents = float(2 * M_PI / numAgents);
        agents.resize(size_t(numAgents));
        for (int i = 0; i < numAgents; ++i) {
            const auto randomRotation = rotationBetweenAgents * i;
            agents[size_t(i)].init(agentSpawnPositions[size_t(i)] + Vec3{0.5f, 0.0f, 0.5f}, randomRotation, lookLimit);
            agents[size_t(i)].updateTransform();
        }
    }
    // addSphere / addDiamond / addPillar (layout_utils.cpp:86-126); returns the index of the created memory object or -1 for the landmark
    void addMemoryShape(int shape, ColorRgb color, Vec3 loc, Vec3 scale, bool collectable, bool good) {
        MemoryObject mo;
        mo.good = good; mo.alive = true;
        mo.root = mul(mat4Translation(loc), mul(mat4Scaling(scale), mat4Identity()));
        const int idx = collectable ? int(memoryObjects.size()) : -1;
        const int pc = paletteIndex(color);
        auto pushDraw = [&](int mesh, int child, const Mat4 &model) {
            if (collectable) drawables[mesh].push_back({DrawEntry::D_MEMORY, idx * 4 + child, pc, mat4Identity()});
            else drawables[mesh].push_back({DrawEntry::D_STATIC, 0, pc, model});
        };
        if (shape == SHAPE_SPHERE) {
            mo.numChildren = 0;
            pushDraw(MESH_SPHERE, 0, mo.root);
        } else if (shape == SHAPE_DIAMOND) {
            mo.numChildren = 1;
            mo.childLocal[0] = mul(mat4Translation({0.0f, -1.0f, 0.0f}), mul(mat4Identity(), mat4RotationX(180.0f * 3.14159265358979323846f / 180.0f)));
            pushDraw(MESH_CONE, 0, mo.root);
            pushDraw(MESH_CONE, 1, mul(mo.root, mo.childLocal[0]));
        } else {
            mo.numChildren = 2;
            const Vec3 capScale{scale.x * 1.2f, 0.15f, scale.z * 1.2f};
            const Vec3 capTranslation = Vec3{0, 0.47f, 0} * scale;
            pushDraw(MESH_CYLINDER, 0, mo.root);
            const Mat4 rootInv = inverted(mo.root);
            for (int k = 0; k < 2; ++k) {  // addCylinder(...)->setParentKeepTransformation(rootObject): local = parentAbs^-1 * abs
                const Vec3 t = k == 0 ? loc + capTranslation : loc - capTranslation;
                const Mat4 capAbs = mul(mat4Translation(t), mul(mat4Scaling(capScale), mat4Identity()));
                mo.childLocal[k] = mul(rootInv, capAbs);
                pushDraw(MESH_CYLINDER, k + 1, mul(mo.root, mo.childLocal[k]));
            }
        }
        if (collectable) memoryObjects.push_back(mo);
    }
    void hexMemoryAddEpisodeDrawables() {  // :156-217
        static const std::vector<int> shapes = {SHAPE_PILLAR, SHAPE_DIAMOND, SHAPE_SPHERE};
        auto goodObjectColor = randomObjectColor(rng), badObjectColor = goodObjectColor;
        auto goodShape = randomSample(shapes, rng), badShape = goodShape;
        while (badObjectColor == goodObjectColor && badShape == goodShape) {
            badObjectColor = randomObjectColor(rng);
            badShape = randomSample(shapes, rng);
        }
        hexMazeAddDrawablesAndCollisions();
        auto shapeScale = [](int shape) {
            if (shape == SHAPE_SPHERE) return Vec3{0.75f, 0.75f, 0.75f};
            if (shape == SHAPE_PILLAR) return Vec3{0.5f, 2, 0.5f};
            return Vec3{0.17f, 0.45f, 0.17f} * float(2.2);
        };
        auto shapeShift = [](int shape) {
            if (shape == SHAPE_SPHERE) return Vec3{0.5f, 0.1f, 0.5f};
            if (shape == SHAPE_PILLAR) return Vec3{0.5f, 0.05f, 0.5f};
            return Vec3{0.5f, 0.6f, 0.5f};
        };
        addMemoryShape(goodShape, goodObjectColor, landmarkLocation + shapeShift(goodShape), shapeScale(goodShape), false, true);
        const float objScale = float(0.6);
        bool isGood = true;
        for (const auto *objects : {&goodObjects, &badObjects}) {
            for (const auto &coord : *objects) {
                const auto shape = isGood ? goodShape : badShape;
                const auto color = isGood ? goodObjectColor : badObjectColor;
                addMemoryShape(shape, color, coord + shapeShift(shape) * objScale, shapeScale(shape) * objScale, true, isGood);
                memoryObjects.back().voxel = vg.grid.getCoords(coord);
            }
            isGood = !isGood;
        }
    }
    void hexMemoryStep() {  // :79-119
        constexpr auto collectRadius = 1.0f;
        if (goodObjectsCollected >= int(goodObjects.size()) && !solved) {
            solved = true;
            doneWithTimer();
        }
        for (int i = 0; i < numAgents; ++i) {
            const Vec3 t = translationOf(agents[size_t(i)].objectT);
            const auto agentCoords = vg.grid.getCoords(t);
            for (int dx = -1; dx <= 1; ++dx)
                for (int dz = -1; dz <= 1; ++dz) {
                    const VoxelCoords coords(agentCoords.x + dx, agentCoords.y, agentCoords.z + dz);
                    // voxel->objects: the collectables whose cell is this voxel, in insertion order
                    for (auto &mo : memoryObjects) {
                        if (!mo.alive || !(mo.voxel == coords)) continue;
                        const Vec3 pillarPos = translationOf(mo.root);
                        const auto distance = length(pillarPos - t);
                        if (distance < collectRadius) {
                            rewardTeam(mo.good ? "memoryCollectGood" : "memoryCollectBad", i, 1);
                            goodObjectsCollected += mo.good;
                            mo.root = mul(mat4Translation({100, 100, 100}), mo.root);
                            mo.alive = false;
                        }
                    }
                }
        }
    }

    // ---------------------------------------------------------------- Sokoban (scenario_sokoban.cpp:83-295)
    void sokobanReloadLevels() {
        const auto levelFilePath = randomSample(allSokobanLevelFiles, rng);
        std::ifstream f{levelFilePath, std::ios::in | std::ios::binary};
        std::string content((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        if (content.empty()) throw std::runtime_error("oracle: could not read " + levelFilePath);
        std::vector<std::string> lines;  // splitString(content, "\n"): strtok_r, i.e. empty lines vanish
        {
            size_t pos = 0;
            while (pos < content.size()) {
                const size_t e = content.find('\n', pos);
                const std::string tok = content.substr(pos, e == std::string::npos ? std::string::npos : e - pos);
                if (!tok.empty()) lines.push_back(tok);
                if (e == std::string::npos) break;
                pos = e + 1;
            }
        }
        SokobanLevel level;
        for (int i = 0; i < int(lines.size()); ++i) {
            if (lines[size_t(i)].find(';') == 0) {
                if (i > 0) sokobanLevels.emplace_back(std::move(level));
                level = SokobanLevel{};
            } else level.rows.emplace_back(lines[size_t(i)]);
        }
        std::shuffle(sokobanLevels.begin(), sokobanLevels.end(), rng);
    }
    void sokobanReset() {
        vg.reset();
        solved = false;
        carryingObject.assign(size_t(numAgents), -1);
        sokobanAgentPositions.clear(), boxesCoords.clear();
        sokoLength = sokoWidth = 0;
        numBoxes = numBoxesOnGoal = 0;
        if (sokobanLevels.empty()) sokobanReloadLevels();
        currLevel = sokobanLevels.back();
        sokobanLevels.pop_back();
        // createLayout (:118-166)
        constexpr int wallHeight = 2;
        const float voxelSize = 2;
        auto &g = vg.grid;
        static const std::vector<ColorRgb> floorColors = {LAYOUT_DEFAULT, VERY_LIGHT_YELLOW, VERY_LIGHT_BLUE, VERY_LIGHT_ORANGE, DARK_GREY};
        auto floorColor = randomSample(floorColors, rng);
        sokoLength = int(currLevel.rows.size());
        for (int x = 0; x < sokoLength; ++x) {
            const auto &row = currLevel.rows[size_t(x)];
            sokoWidth = std::max(sokoWidth, int(row.size()));
            for (int z = 0; z < int(row.size()); ++z) {
                g.set({x, 0, z}, VoxelGridComponent::makeVoxel(VOXEL_SOLID | VOXEL_OPAQUE, TERRAIN_NONE, floorColor));
                if (row[size_t(z)] == '#') {
                    for (int y = 1; y <= wallHeight; ++y) g.set({x, y, z}, VoxelGridComponent::makeVoxel(VOXEL_SOLID));
                    g.get(VoxelCoords{x, 1, z})->terrain = SOKO_WALL;
                }
                if (row[size_t(z)] == '@' || row[size_t(z)] == '+') {
                    for (int agentIdx = 0; agentIdx < numAgents; ++agentIdx) {
                        const float agentX = float(x) + float(agentIdx % 2) * 0.5f;
                        const float agentZ = float(z) + float(agentIdx % 4 > 1) * 0.5f;
                        sokobanAgentPositions.emplace_back(Vec3{agentX * voxelSize, float(voxelSize + 0.3 * float(agentIdx) * voxelSize), agentZ * voxelSize});
                    }
                }
                if (row[size_t(z)] == '.' || row[size_t(z)] == '+') g.set({x, 1, z}, VoxelGridComponent::makeVoxel(VOXEL_EMPTY, SOKO_GOAL));
                if (row[size_t(z)] == '$' || row[size_t(z)] == '*') { boxesCoords.emplace_back(x, 1, z); ++numBoxes; }
            }
        }
        agentSpawnPositions = sokobanAgentPositions;
        agentSpawnPositions.resize(size_t(numAgents), Vec3{0, 0, 0});  // a level without a player cell would index past the end upstream
        agentInitialPositions = agentSpawnPositions;
        objectSpawnPositions = boxesCoords; rewardSpawnPositions.clear();
    }
    void sokobanAddEpisodeDrawables() {  // :229-295
        const float voxelSize = 2;
        addDrawablesAndCollisionObjectsFromVoxelGrid(vg.grid.getVoxelSize());
        auto &g = vg.grid;
        for (int x = 0; x < sokoLength; ++x)
            for (int z = 0; z < sokoWidth; ++z) {
                if (!g.hasVoxel({x, 1, z})) continue;
                auto v = g.get({x, 1, z});
                if (v->terrain == SOKO_EMPTY) continue;
                const Vec3 pos{voxelSize * float(x) + voxelSize / 2, voxelSize, voxelSize * float(z) + voxelSize / 2};
                const float h = v->terrain == SOKO_WALL ? 0.35f : 0.025f;
                Mat4 m = mul(mat4Scaling({1, h, 1}), mat4Identity());
                m = mul(mat4Translation({0.0f, h, 0.0f}), m);
                m = mul(mat4Translation(pos), m);
                drawables[MESH_BOX].push_back({DrawEntry::D_STATIC, 0, paletteIndex(v->terrain == SOKO_WALL ? LIGHT_ORANGE : LIGHT_GREEN), m});
            }
        for (auto box : boxesCoords) {
            const Vec3 scale = Vec3{voxelSize / 2, 0.45f, voxelSize / 2} * 0.8f;
            const Vec3 translation = Vec3{float(box.x) + 0.5f, float(box.y) + 0.2f, float(box.z) + 0.5f} * voxelSize;
            MovableObject o;
            o.local = mul(mat4Translation(translation), mul(mat4Scaling(scale), mat4Identity()));
            o.collisionScale = Vec3{1.15f, 3, 1.15f};
            o.collisionOffset = Vec3{0, 0.6f, 0};
            o.color = paletteIndex(DARK_BLUE);
            o.collider = int(colliders.size());
            colliders.push_back(Collider{});
            objects.push_back(o);
            const int idx = int(objects.size()) - 1;
            drawables[MESH_BOX].push_back({DrawEntry::D_OBJECT, idx, o.color, mat4Identity()});
            syncPose(idx);
            if (!g.hasVoxel(box)) g.set(box, VoxelGridComponent::makeVoxel(VOXEL_EMPTY));
            g.get(box)->physicsObject = idx;
        }
    }
    void sokobanStep() {  // :168-222
        const float voxelSize = 2;
        for (int i = 0; i < numAgents; ++i) {
            const auto a = currAction[size_t(i)];
            Agent &agent = agents[size_t(i)];
            if (!(a & A_Interact)) continue;
            const Vec3 t = translationOf(agent.pickupAbs());
            auto voxel = vg.grid.getWithVector(t);
            if (voxel && voxel->physicsObject >= 0) {
                const auto boxPos = vg.grid.getCoords(t);
                const auto agentPos = vg.grid.getCoords(translationOf(agent.objectT));
                const int dist = std::abs(agentPos.x - boxPos.x) + std::abs(agentPos.y - boxPos.y) + std::abs(agentPos.z - boxPos.z);
                if (dist == 1) {
                    const VoxelCoords deltaPos(boxPos.x - agentPos.x, boxPos.y - agentPos.y, boxPos.z - agentPos.z);
                    const VoxelCoords desiredPos(boxPos.x + deltaPos.x, boxPos.y + deltaPos.y, boxPos.z + deltaPos.z);
                    bool occupied = false;
                    for (int j = 0; j < numAgents; ++j)
                        if (vg.grid.getCoords(translationOf(agents[size_t(j)].objectT)) == desiredPos) { occupied = true; break; }
                    if (!occupied) {
                        if (!vg.grid.hasVoxel(desiredPos)) vg.grid.set(desiredPos, VoxelGridComponent::makeVoxel(VOXEL_EMPTY));
                        voxel = vg.grid.get(boxPos);  // (the insertion above may rehash the map)
                        auto desiredPosVoxel = vg.grid.get(desiredPos);
                        if (desiredPosVoxel->terrain != SOKO_WALL && desiredPosVoxel->physicsObject < 0) {
                            const int obj = voxel->physicsObject;
                            desiredPosVoxel->physicsObject = obj;
                            objects[size_t(obj)].local = mul(mat4Translation(Vec3{float(deltaPos.x), float(deltaPos.y), float(deltaPos.z)} * voxelSize), objects[size_t(obj)].local);
                            syncPose(obj);
                            voxel->physicsObject = -1;
                            if (voxel->terrain != SOKO_GOAL && desiredPosVoxel->terrain == SOKO_GOAL) {
                                ++numBoxesOnGoal;
                                rewardTeam("sokobanBoxOnTarget", i, 1);
                                if (numBoxesOnGoal == numBoxes && !solved) {
                                    solved = true;
                                    rewardTeam("sokobanAllBoxesOnTarget", i, 1);
                                    doneWithTimer();
                                }
                            } else if (voxel->terrain == SOKO_GOAL && desiredPosVoxel->terrain != SOKO_GOAL) {
                                --numBoxesOnGoal;
                                rewardTeam("sokobanBoxLeavesTarget", i, 1);
                            }
                        }
                    }
                }
            }
        }
    }

    // ---------------------------------------------------------------- Rearrange (scenario_rearrange.cpp:46-300)
    static ArrangementItem randomArrangementItem(Rng &rng, const VoxelCoords &offset) {  // scenario_rearrange.hpp:31-46
        static const std::vector<int> shapes{MESH_CYLINDER, MESH_CAPSULE, MESH_BOX, MESH_SPHERE};
        ArrangementItem item;
        item.shape = randomSample(shapes, rng);
        item.color = randomObjectColor(rng);
        item.offset = offset;
        return item;
    }
    void rearrangeReset() {
        solved = false;
        vg.reset();
        carryingObject.assign(size_t(numAgents), -1);
        levelRoot = std::make_unique<Node>();
        rearrangePlatform = std::make_unique<RearrangePlatform>(levelRoot.get(), rng, WALLS_ALL, floatParams);
        rearrangePlatform->init(), rearrangePlatform->generate();
        vg.addPlatform(*rearrangePlatform, DARK_GREY, DARK_GREY, randomBool(rng));
        arrangement.clear();
        arrangementObjects.clear();
        generateArrangement();
        // DefaultScenario::spawnAgents asks the scenario for the positions first (scenario_default.hpp:80-97)
        agentSpawnPositions.assign(size_t(numAgents), Vec3{0, 0, 0});
        for (int i = 0; i < numAgents; ++i)
            for (int attempt = 0; attempt < 20; ++attempt) {  // :179-200
                int agentX = randRange(2, rearrangePlatform->length - 1, rng);
                int agentZ = randRange(2, rearrangePlatform->width - 1, rng);
                if (std::fabs(agentX - leftCenter.x) < 2 && std::fabs(agentZ - leftCenter.z) < 2) continue;
                if (std::fabs(agentX - rightCenter.x) < 2 && std::fabs(agentZ - rightCenter.z) < 2) continue;
                agentSpawnPositions[size_t(i)] = Vec3{float(agentX), 2, float(agentZ)};
                break;
            }
        agentInitialPositions = agentSpawnPositions;
        objectSpawnPositions.clear(); rewardSpawnPositions.clear();
    }
    void generateArrangement() {  // :70-126
        const int arrangementSize = randRange(2, 8, rng);
        std::queue<ArrangementItem> q;
        std::unordered_set<VoxelCoords, VoxelHash> used;
        const auto firstItem = randomArrangementItem(rng, {0, 0, 0});
        q.push(firstItem);
        arrangement.emplace_back(firstItem);
        used.insert({0, 0, 0});
        std::vector<VoxelCoords> directions{{-1, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};
        while (!q.empty()) {
            const auto currItem = q.front();
            q.pop();
            int maxBranches = randRange(1, int(directions.size()) + 1, rng);
            maxBranches = randRange(1, maxBranches + 1, rng);
            int numBranches = 0;
            std::shuffle(directions.begin(), directions.end(), rng);
            for (auto dir : directions) {
                const VoxelCoords newOffset(currItem.offset.x + dir.x, currItem.offset.y + dir.y, currItem.offset.z + dir.z);
                const VoxelCoords below(newOffset.x, newOffset.y - 1, newOffset.z);
                if (newOffset.y >= 2 || std::abs(newOffset.x) >= 2 || std::abs(newOffset.z) >= 2) continue;
                if (used.count(newOffset)) continue;
                if (!(newOffset.y == 0 || used.count(below))) continue;
                const auto newItem = randomArrangementItem(rng, newOffset);
                q.push(newItem);
                arrangement.emplace_back(newItem);
                used.insert(newOffset);
                ++numBranches;
                if (numBranches >= maxBranches) break;
                if (int(arrangement.size()) >= arrangementSize) break;
            }
            if (int(arrangement.size()) >= arrangementSize) break;
        }
    }
    bool arrangementContains(int shape, ColorRgb color, const VoxelCoords &offset) const {  // scenario_rearrange.hpp:77-84
        for (const auto &item : arrangement)
            if (item.shape == shape && item.color == color && item.offset == offset) return true;
        return false;
    }
    int countMatchingObjects() const {  // :134-149
        int matching = 0;
        for (int obj : arrangementObjects) {
            if (objects[size_t(obj)].pickedUp) continue;
            const VoxelCoords voxelCoords = vg.grid.getCoords(translationOf(objectAbs(obj)));
            const VoxelCoords offset(voxelCoords.x - rightCenter.x, voxelCoords.y - rightCenter.y, voxelCoords.z - rightCenter.z);
            if (arrangementContains(objects[size_t(obj)].mesh, objects[size_t(obj)].arrangementColor, offset)) ++matching;
        }
        return matching;
    }
    void rearrangeCheckDone(int agentIdx) {  // :163-177
        const auto matches = countMatchingObjects();
        if (matches > maxMatchingObjects) {
            rewardTeam("rearrangeOneMoreObjectCorrectPosition", agentIdx, 1);
            maxMatchingObjects = matches;
        }
        if (matches >= int(arrangement.size()) && !solved) {
            solved = true;
            rewardTeam("rearrangeAllObjectsCorrectPosition", agentIdx, 1);
            doneWithTimer();
        }
    }
    void arrangementDrawables(VoxelCoords center, bool interactive) {  // :202-263
        auto shapeScale = [](int shape) {
            if (shape == MESH_CAPSULE) return Vec3{0.8f, 0.5f, 0.8f};
            if (shape == MESH_CYLINDER) return Vec3{0.9f, 2, 0.9f};
            return Vec3{1, 1, 1};
        };
        const auto objSize = 0.45f;
        std::unordered_set<VoxelCoords, VoxelHash> occupied;
        for (const auto &item : arrangement) occupied.insert(item.offset);
        int numUnmovedItems = int(arrangement.size());
        if (interactive) numUnmovedItems = randRange(0, int(arrangement.size()), rng);
        int placedItems = 0;
        for (const auto &item : arrangement) {
            VoxelCoords pos(item.offset.x + center.x, item.offset.y + center.y, item.offset.z + center.z);
            if (interactive && placedItems >= numUnmovedItems) {
                VoxelCoords newOffset = item.offset;
                while (occupied.count(newOffset)) {
                    const int rx = randRange(-2, 3, rng);  // braced initialiser: left to right
                    const int rz = randRange(-2, 3, rng);
                    newOffset = VoxelCoords(rx, 0, rz);
                }
                pos = VoxelCoords(newOffset.x + center.x, newOffset.y + center.y, newOffset.z + center.z);
                occupied.insert(newOffset);
            }
            const Vec3 translation{float(pos.x) + 0.5f, float(pos.y) + 0.5f, float(pos.z) + 0.5f};
            MovableObject o;
            o.local = mul(mat4Translation(translation), mul(mat4Scaling(shapeScale(item.shape) * objSize), mat4Identity()));
            o.collisionScale = item.shape == MESH_CYLINDER ? Vec3{1, 0.5f, 1} : (item.shape == MESH_CAPSULE ? Vec3{1, 2, 1} : Vec3{1, 1, 1});
            o.collisionOffset = Vec3{0, 0, 0};
            o.color = paletteIndex(item.color);
            o.arrangementColor = item.color;
            o.mesh = item.shape;
            o.collider = int(colliders.size());
            colliders.push_back(Collider{});
            if (interactive) {
                objects.push_back(o);
                const int idx = int(objects.size()) - 1;
                syncPose(idx);
                drawables[item.shape].push_back({DrawEntry::D_OBJECT, idx, o.color, mat4Identity()});
                Voxel voxelState;
                voxelState.physicsObject = idx;
                vg.grid.set(pos, voxelState);
                arrangementObjects.emplace_back(idx);
                objectSpawnPositions.push_back(pos);  // (introspection only: Rearrange's episode length does not use it)
            } else {  // never picked up: a static collider + drawable
                Collider &c = colliders.back();
                c.kind = 0;
                c.c = translationOf(o.local) + o.collisionOffset;
                c.h = scalingOf(o.local) * o.collisionScale;
                drawables[item.shape].push_back({DrawEntry::D_STATIC, 0, o.color, o.local});
            }
            ++placedItems;
        }
    }
    void rearrangeAddEpisodeDrawables() {  // :265-300
        addDrawablesAndCollisionObjectsFromVoxelGrid(1.0f);
        for (int dx = -3; dx <= 3; ++dx)
            for (int dz = -3; dz <= 3; ++dz) {
                Voxel voxelState;
                voxelState.voxelType = VOXEL_SOLID;
                vg.grid.set(VoxelCoords(leftCenter.x + dx, 1, leftCenter.z + dz), voxelState);
                vg.grid.set(VoxelCoords(rightCenter.x + dx, 1, rightCenter.z + dz), voxelState);
            }
        arrangementDrawables(leftCenter, false);
        arrangementDrawables(rightCenter, true);
        maxMatchingObjects = countMatchingObjects();
        const Vec3 platformCenter{9.5f, 0, 7};
        const Vec3 lc{float(leftCenter.x), float(leftCenter.y), float(leftCenter.z)}, rc{float(rightCenter.x), float(rightCenter.y), float(rightCenter.z)};
        addStaticCollidingBox({8.35f, 0.5f, 5.65f}, platformCenter + Vec3{0.0f, 1, 0.0f}, DARK_GREY);
        addStaticCollidingBox({3, 0.5f, 3}, lc + Vec3{0.5f, -0.5f, 0.5f}, LAYOUT_DEFAULT);
        addStaticCollidingBox({1.5f, 0.5f, 1.5f}, lc + Vec3{0.5f, -0.45f, 0.5f}, DARK_GREY);
        addStaticCollidingBox({3, 0.5f, 3}, lc + Vec3{1.0f, -0.66f, 1.0f}, LAYOUT_DEFAULT);
        addStaticCollidingBox({3, 0.5f, 3}, lc + Vec3{1.5f, -0.82f, 1.5f}, LAYOUT_DEFAULT);
        addStaticCollidingBox({3, 0.5f, 3}, rc + Vec3{0.5f, -0.5f, 0.5f}, BLUE);
        addStaticCollidingBox({1.5f, 0.5f, 1.5f}, rc + Vec3{0.5f, -0.45f, 0.5f}, DARK_GREY);
        addStaticCollidingBox({3, 0.5f, 3}, rc + Vec3{0, -0.66f, 1.0f}, BLUE);
        addStaticCollidingBox({3, 0.5f, 3}, rc + Vec3{-0.5f, -0.82f, 1.5f}, BLUE);
    }
    void rearrangeStep() {  // ObjectStackingComponent::step (component_object_stacking.hpp:45-56)
        for (int i = 0; i < numAgents; ++i)
            if (currAction[i] & A_Interact) onInteractAction(i);
    }

    // ---------------------------------------------------------------- Collect (scenario_collect.cpp:20-218)
    void collectReset() {
        solved = false;
        vg.reset();
        carryingObject.assign(size_t(numAgents), -1);
        numPositiveRewards = positiveRewardsCollected = 0;
        agentSpawnPositions.clear(); objectSpawnPositions.clear(); rewardSpawnPositions.clear();
        static const std::vector<ColorRgb> landscapeColors = {LAYOUT_DEFAULT, VERY_LIGHT_GREEN, VERY_LIGHT_BLUE, VERY_LIGHT_GREY, VERY_LIGHT_ORANGE, GREY, DARK_GREY};
        static const std::vector<ColorRgb> floorColors = {GREY, DARK_GREY, DARK_GREY};
        auto landscapeColor = randomSample(landscapeColors, rng);
        auto floorColor = randomSample(floorColors, rng);
        constexpr int maxWidth = 42, maxLength = maxWidth;
        const int width = randRange(8, maxWidth, rng);
        const int length = randRange(8, maxWidth, rng);
        std::vector<int> spawnHeight(size_t(length * width), 1);
        double frequency = double(randRange(1, 100, rng)) / 10.0;
        const std::int32_t octaves = randRange(1, 10, rng);
        const std::uint32_t seed = randRange(0, 1000000000, rng);
        const PerlinNoise perlin(seed);
        const double fx = maxLength / frequency;
        const double fz = maxWidth / frequency;
        const int intensity = randRange(5, 18, rng);
        const float groundLevel = frand(rng) * 0.5f + 0.2f;
        for (int x = 1; x < length - 1; ++x)
            for (int z = 1; z < width - 1; ++z) {
                const double noise = perlin.accumulatedOctaveNoise2D_0_1(x / fx, z / fz, octaves);
                const double yCoord = intensity * (noise - groundLevel);
                if (yCoord >= 1) {
                    const int yCoordRound = int(lround(yCoord));
                    for (int y = yCoordRound; y >= 1; --y) vg.grid.set({x, y, z}, VoxelGridComponent::makeVoxel(VOXEL_SOLID | VOXEL_OPAQUE, TERRAIN_NONE, landscapeColor));
                    spawnHeight[size_t(x * width + z)] = yCoordRound + 1;
                }
            }
        for (int x = 0; x < length; ++x)
            for (int z = 0; z < width; ++z) vg.grid.set({x, 0, z}, VoxelGridComponent::makeVoxel(VOXEL_SOLID | VOXEL_OPAQUE, TERRAIN_NONE, floorColor));
        std::vector<VoxelCoords> spawnPositions;
        for (int x = 1; x < length - 1; ++x)
            for (int z = 1; z < width - 1; ++z) spawnPositions.emplace_back(x, spawnHeight[size_t(x * width + z)], z);
        std::shuffle(spawnPositions.begin(), spawnPositions.end(), rng);
        int offset = 0;
        for (int i = 0; i < numAgents; ++i) agentSpawnPositions.emplace_back(float(spawnPositions[size_t(i)].x), float(spawnPositions[size_t(i)].y), float(spawnPositions[size_t(i)].z));
        offset += numAgents;
        int numRewards = randRange(1, int(lround(0.05 * width * length)) + 2, rng);
        numRewards = std::min(numRewards, int(spawnPositions.size()) - offset);
        int numRewardsPlacedRandomly = std::max(numRewards / 2, 1);
        rewardSpawnPositions = std::vector<VoxelCoords>(spawnPositions.begin() + offset, spawnPositions.begin() + offset + numRewardsPlacedRandomly);
        offset += numRewardsPlacedRandomly;
        std::sort(spawnPositions.begin() + offset, spawnPositions.end(), [&](const VoxelCoords &a, const VoxelCoords &b) {
            int heightA = spawnHeight[size_t(a.x * width + a.z)];
            int heightB = spawnHeight[size_t(b.x * width + b.z)];
            if (heightA != heightB) return heightA > heightB;
            else return false;
        });
        rewardSpawnPositions.insert(rewardSpawnPositions.end(), spawnPositions.begin() + offset, spawnPositions.begin() + offset + (numRewards - numRewardsPlacedRandomly));
        offset += numRewards - numRewardsPlacedRandomly;
        std::shuffle(spawnPositions.begin() + offset, spawnPositions.end(), rng);
        auto objectsMin = std::max(3, int(length * width * 0.04));
        auto objectsMax = std::min(objectsMin + 1, int(lround(0.07 * width * length)) + 2);
        const int numObjects = std::min(randRange(objectsMin, objectsMax, rng), int(spawnPositions.size()) - offset);
        if (offset + numObjects < int(spawnPositions.size())) {
            objectSpawnPositions = std::vector<VoxelCoords>(spawnPositions.begin() + offset, spawnPositions.begin() + offset + numObjects);
            offset += numObjects;
        }
        agentInitialPositions = agentSpawnPositions;
    }

    void collectAddEpisodeDrawables() {  // scenario_collect.cpp:185-212
        addDrawablesAndCollisionObjectsFromVoxelGrid(1.0f);
        addObjects(objectSpawnPositions);
        for (const auto &pos : rewardSpawnPositions) {
            RewardObject ro;
            const Vec3 translation{float(pos.x) + 0.5f, float(pos.y) + 0.8f, float(pos.z) + 0.5f};
            Voxel voxel;
            if (frand(rng) > 0.3f) { voxel.reward = +1; ro.color = paletteIndex(GREEN); ++numPositiveRewards; }
            else { voxel.reward = -1; ro.color = paletteIndex(RED); }
            ro.bottomLocal = mul(mat4Translation({0.0f, -1.0f, 0.0f}), mul(mat4Identity(), mat4RotationX(180.0f * 3.14159265358979323846f / 180.0f)));
            ro.root = mul(mat4Translation(translation), mul(mat4Scaling({0.17f, 0.45f, 0.17f}), mat4Identity()));
            voxel.rewardObject = int(rewardObjects.size());
            addRewardDrawables(ro);
            rewardObjects.push_back(ro);
            vg.grid.set(pos, voxel);
        }
    }

    void collectStep() {  // scenario_collect.cpp:145-178,214-218
        for (int i = 0; i < numAgents; ++i)
            if (currAction[i] & A_Interact) onInteractAction(i);
        for (int i = 0; i < numAgents; ++i)  // FallDetectionComponent::step with the agentFell callback
            if (translationOf(agents[i].objectT).y < -20) { resetAgent(i); rewardAgent("collectSingleBad", i, 1); }
        for (int i = 0; i < numAgents; ++i) {
            const Vec3 t = translationOf(agents[i].objectT);
            VoxelCoords voxel = vg.grid.getCoords(t);
            auto voxelPtr = vg.grid.get(voxel);
            if (voxelPtr && voxelPtr->rewardObject >= 0) {
                RewardObject &ro = rewardObjects[size_t(voxelPtr->rewardObject)];
                ro.root = mul(mat4Translation({500, 500, 500}), ro.root);
                if (voxelPtr->reward > 0) ++positiveRewardsCollected;
                if (voxelPtr->reward > 0) rewardTeam("collectSingleGood", i, 1);
                else if (voxelPtr->reward < 0) rewardTeam("collectSingleBad", i, 1);
                if (positiveRewardsCollected >= numPositiveRewards && !solved) {
                    solved = true;
                    doneWithTimer();
                    rewardTeam("collectAll", i, 1);
                }
                vg.grid.remove(voxel);
            }
        }
    }

    // ---------------------------------------------------------------- Obstacles (scenario_obstacles.cpp:51-278)
    std::unique_ptr<Platform> makePlatform(Node *parent, int walls, int width) {  // :17-37
        const auto platformType = randomSample(platformTypes, rng);
        switch (platformType) {
            case PT_STEP: return std::make_unique<StepPlatform>(parent, rng, walls, floatParams, width);
            case PT_GAP: return std::make_unique<GapPlatform>(parent, rng, walls, floatParams, width);
            case PT_LAVA: return std::make_unique<LavaPlatform>(parent, rng, walls, floatParams, width);
            case PT_WALL: return std::make_unique<WallPlatform>(parent, rng, walls, floatParams, width);
            default: return std::make_unique<EmptyPlatform>(parent, rng, walls, floatParams, width);
        }
    }

    void obstaclesReset() {
        vg.reset();
        platforms.clear();
        levelRoot = std::make_unique<Node>();
        carryingObject.assign(size_t(numAgents), -1);
        agentSpawnPositions.clear(); objectSpawnPositions.clear(); rewardSpawnPositions.clear();
        agentReachedExit.assign(size_t(numAgents), false);
        solved = false;

        const bool drawWalls = randRange(0, 2, rng);
        Platform *startPlatform = nullptr;
        for (int attempt = 0; attempt < 20; ++attempt) {
            platforms.clear();
            numPlatforms = randRange(int(lroundf(floatParams["obstaclesMinNumPlatforms"])), int(lroundf(floatParams["obstaclesMaxNumPlatforms"])) + 1, rng);
            static const std::vector<int> orientations = {0, 1, 2};  // STRAIGHT, TURN_LEFT, TURN_RIGHT
            auto startPlatformPtr = std::make_unique<StartPlatform>(levelRoot.get(), rng, floatParams);
            startPlatformPtr->init(), startPlatformPtr->generate();
            int requiredWidth = startPlatformPtr->width;
            startPlatform = startPlatformPtr.get();
            Platform *previousPlatform = startPlatform;
            platforms.emplace_back(std::move(startPlatformPtr));
            int numMaxDifficultyObstacles = 0;
            const int numAllowedMaxDifficultyObstacles = int(floatParams.at("obstaclesNumAllowedMaxDifficulty"));
            for (int i = 0; i < numPlatforms; ++i) {
                auto orientation = randomSample(orientations, rng);
                requiredWidth = orientation == 0 ? requiredWidth : -1;
                std::unique_ptr<Platform> newPlatform;
                while (!newPlatform || (newPlatform->isMaxDifficulty() && numMaxDifficultyObstacles >= numAllowedMaxDifficultyObstacles)) {
                    newPlatform = makePlatform(previousPlatform->nextPlatformAnchor, WALLS_WEST | WALLS_EAST, requiredWidth);
                    newPlatform->init();
                }
                if (newPlatform->isMaxDifficulty()) ++numMaxDifficultyObstacles;
                platforms.emplace_back(std::move(newPlatform));
                auto platform = platforms.back().get();
                platform->generate();
                if (orientation == 1) platform->rotateCCW(previousPlatform->width);
                else if (orientation == 2) platform->rotateCW(previousPlatform->width);
                if (orientation != 0) {
                    int walls = WALLS_NORTH;
                    walls |= orientation == 1 ? WALLS_WEST : WALLS_EAST;
                    const int w = previousPlatform->width, l = platform->width - 1;
                    platforms.emplace_back(std::make_unique<TransitionPlatform>(previousPlatform->nextPlatformAnchor, rng, walls, floatParams, l, w));
                    auto transitionPlatform = platforms.back().get();
                    transitionPlatform->init();
                    transitionPlatform->generate();
                }
                previousPlatform = platform;
                requiredWidth = platform->width;
            }
            auto exitPlatformPtr = std::make_unique<ExitPlatform>(previousPlatform->nextPlatformAnchor, rng, floatParams, requiredWidth);
            exitPlatformPtr->init(), exitPlatformPtr->generate();
            platforms.emplace_back(std::move(exitPlatformPtr));
            bool selfCollision = false;
            for (int j = 0; j < int(platforms.size()) && !selfCollision; ++j)
                for (int k = 0; k < j - 2; ++k)
                    if (platforms[j]->collidesWith(*platforms[k])) { selfCollision = true; break; }
            if (!selfCollision) break;
        }
        auto layoutColor = randomLayoutColor(rng);
        auto wallColor = randomLayoutColor(rng);
        for (auto &p : platforms) vg.addPlatform(*p, layoutColor, wallColor, drawWalls);

        agentSpawnPositions = startPlatform->agentSpawnPoints(numAgents);
        // ten random attempts per agent (platforms.hpp:221-244) can leave fewer spawn points than agents on a small start platform; the
        // reference then indexes past the end of this vector in spawnAgents (scenario_default.hpp:83-91) -- undefined behaviour, whatever
        // the heap holds.  The oracle makes it defined (the origin, what a zeroed heap gives) and remembers; the product refuses such
        // configurations ("start platform too small for the agents"), and the parity tests stay away from them.
        if (int(agentSpawnPositions.size()) < numAgents) { undefinedSpawn = true; agentSpawnPositions.resize(size_t(numAgents), Vec3{0, 0, 0}); }
        agentInitialPositions = agentSpawnPositions;

        std::vector<int> numBoxes(platforms.size());
        for (int i = 1; i < int(platforms.size()); ++i) {
            const auto n = platforms[i]->requiresMovableBoxesToTraverse();
            for (int box = 0; box < n; ++box) {
                const auto platformIdx = randRange(std::max(0, i - 2), i, rng);
                ++numBoxes[platformIdx];
            }
        }
        for (int i = 0; i < int(platforms.size()); ++i) {
            float randomBoxesFraction = frand(rng) * 0.5f;
            auto randomBoxes = int(lroundf(randomBoxesFraction * numBoxes[i])) + randRange(0, 2, rng);
            const auto coords = platforms[i]->generateObjectPositions(numBoxes[i] + randomBoxes);
            objectSpawnPositions.insert(objectSpawnPositions.end(), coords.cbegin(), coords.cend());
        }
        for (int i = 1; i < int(platforms.size()) - 1; ++i) {
            auto numRewardObjects = randRange(0, 2, rng);
            const auto coords = platforms[i]->generateObjectPositions(numRewardObjects);
            rewardSpawnPositions.insert(rewardSpawnPositions.end(), coords.cbegin(), coords.cend());
        }
    }

    void obstaclesAddEpisodeDrawables() {  // scenario_obstacles.cpp:241-260
        addDrawablesAndCollisionObjectsFromVoxelGrid(1.0f);
        for (auto &platform : platforms)
            for (auto &[terrainType, boxes] : platform->terrainBoxes)
                for (auto &bb : boxes) addTerrain(terrainType, bb.boundingBox());
        addObjects(objectSpawnPositions);
        for (const auto &pos : rewardSpawnPositions) {
            // addDiamond (layout_utils.cpp:114-126): two cones, the lower one rotated 180 deg about X and shifted -1 in y
            RewardObject ro;
            const Vec3 translation = Vec3{float(pos.x), float(pos.y), float(pos.z)} + Vec3{0.5f, 0.7f, 0.5f};
            const Vec3 scale = Vec3{0.17f, 0.45f, 0.17f} * 0.8f;
            ro.bottomLocal = mul(mat4Translation({0.0f, -1.0f, 0.0f}), mul(mat4Identity(), mat4RotationX(180.0f * 3.14159265358979323846f / 180.0f)));
            ro.root = mul(mat4Translation(translation), mul(mat4Scaling(scale), mat4Identity()));
            ro.color = paletteIndex(GREEN);
            if (!vg.grid.hasVoxel(pos)) vg.grid.set(pos, Voxel{});
            vg.grid.get(pos)->rewardObject = int(rewardObjects.size());
            addRewardDrawables(ro);
            rewardObjects.push_back(ro);
        }
    }

    void obstaclesStep() {  // scenario_obstacles.cpp:197-239
        for (int i = 0; i < numAgents; ++i)
            if (currAction[i] & A_Interact) onInteractAction(i);
        fallDetectionStep();
        int numAgentsAtExit = 0;
        for (int i = 0; i < numAgents; ++i) {
            const Vec3 t = translationOf(agents[i].objectT);
            const auto voxel = vg.grid.getCoords(t);
            if (vg.grid.hasVoxel(voxel)) {
                const auto terrainType = vg.grid.get(voxel)->terrain;
                if (terrainType & TERRAIN_EXIT) {
                    ++numAgentsAtExit;
                    if (!agentReachedExit[i]) {
                        agentReachedExit[i] = true;
                        rewardTeam("obstaclesAgentAtExit", i, 1);
                        if (carryingObject[i] >= 0) rewardTeam("obstaclesAgentCarriedObjectToExit", i, 1);
                    }
                } else if (terrainType & TERRAIN_LAVA)
                    resetAgent(i);  // agentTouchedLava -> fallDetection.resetAgent; agentFell() gives no penalty
                auto voxelData = vg.grid.get(voxel);
                if (voxelData->rewardObject >= 0) {
                    RewardObject &ro = rewardObjects[voxelData->rewardObject];
                    ro.root = mul(mat4Translation({1000, 1000, 1000}), ro.root);
                    voxelData->rewardObject = -1;
                    rewardTeam("obstaclesExtraReward", i, 1);
                }
            }
        }
        if (numAgentsAtExit == numAgents && !solved) {
            solved = true;
            doneWithTimer();
            for (int i = 0; i < numAgents; ++i) rewardAgent("obstaclesAllAgentsAtExit", i, 1);  // rewardAll
        }
    }

    // TowerBuildingPlatform (scenario_tower_building.cpp:8-115)
    struct TowerPlatform : EmptyPlatform {
        TowerPlatform(Node *parent, Rng &rng, int walls, const FloatParams &p, int numAgents) : EmptyPlatform(parent, rng, walls, p), numAgents(numAgents) {}
        void init() override {
            height = randRange(5, 7, rng);
            length = randRange(12, 30, rng);
            width = randRange(12, 25, rng);
            buildZoneLength = randRange(3, 9, rng);
            buildZoneWidth = randRange(3, 9, rng);
            materialsLength = randRange(2, 8, rng);
            materialsWidth = randRange(2, 8, rng);
            length = std::max(buildZoneLength + materialsLength + 3, length);
            width = std::max(buildZoneWidth + materialsWidth + 3, width);
            buildZoneXOffset = randRange(1, length - buildZoneLength - 1, rng);
            buildZoneZOffset = randRange(1, width - buildZoneWidth - 1, rng);
            materialsXOffset = randRange(1, length - materialsLength - 1, rng);
            materialsZOffset = randRange(1, width - materialsWidth - 1, rng);
            std::vector<VoxelCoords> spawnCandidates;
            for (int x = 1; x < length - 1; ++x)
                for (int z = 1; z < width - 1; ++z) spawnCandidates.emplace_back(x, 2, z);
            std::shuffle(spawnCandidates.begin(), spawnCandidates.end(), rng);
            agentSpawnCoords.clear();
            for (int i = 0; i < std::min(numAgents, int(spawnCandidates.size())); ++i)
                agentSpawnCoords.emplace_back(float(spawnCandidates[i].x), float(spawnCandidates[i].y), float(spawnCandidates[i].z));
            auto spawnIdx = int(agentSpawnCoords.size());
            const auto maxRandomObjects = std::min(int(spawnCandidates.size()) - numAgents, 25);
            const auto spawnObjects = randRange(0, std::max(1, maxRandomObjects), rng);
            objectSpawnCoords = std::vector<VoxelCoords>(spawnCandidates.begin() + spawnIdx, spawnCandidates.begin() + spawnIdx + spawnObjects);
            for (auto &c : objectSpawnCoords) {
                if (c.x >= materialsXOffset && c.x < materialsXOffset + materialsLength && c.z >= materialsZOffset && c.z < materialsZOffset + materialsWidth) continue;
                c.y -= 1;
            }
            for (int x = materialsXOffset; x < materialsXOffset + materialsLength; ++x)
                for (int y = 1; y <= 1; ++y)
                    for (int z = materialsZOffset; z < materialsZOffset + materialsWidth; ++z) objectSpawnCoords.emplace_back(x, y, z);
            while (int(agentSpawnCoords.size()) < numAgents) agentSpawnCoords.emplace_back(agentSpawnCoords[0]);
        }
        void generate() override {
            EmptyPlatform::generate();
            const VoxelCoords minCoord{buildZoneXOffset, 1, buildZoneZOffset}, maxCoord{buildZoneXOffset + buildZoneLength, 1, buildZoneZOffset + buildZoneWidth};
            terrainBoxes[TERRAIN_BUILDING_ZONE].emplace_back(makeAABB({minCoord, maxCoord}));
        }
        std::vector<Vec3> agentSpawnCoords;
        std::vector<VoxelCoords> objectSpawnCoords;
        int buildZoneLength{}, buildZoneWidth{}, materialsLength{}, materialsWidth{};
        int buildZoneXOffset{}, buildZoneZOffset{}, materialsXOffset{}, materialsZOffset{};
        int numAgents{};
    };

    void towerReset() {  // scenario_tower_building.cpp:129-154
        std::fill(carryingObject.begin(), carryingObject.end(), -1);
        carryingObject.assign(size_t(numAgents), -1);
        vg.reset();
        agentInitialPositions.clear();
        levelRoot = std::make_unique<Node>();
        std::fill(agentState.begin(), agentState.end(), TowerAgentState{});

        auto layoutColor = randomLayoutColor(rng);
        while (layoutColor == BUILDING_ZONE) layoutColor = randomLayoutColor(rng);

        platform = std::make_unique<TowerPlatform>(levelRoot.get(), rng, WALLS_ALL, floatParams, numAgents);
        platform->init(), platform->generate();
        // GCC evaluates call arguments right to left: randomBool (drawWalls) BEFORE randomLayoutColor (wall colour)
        const bool drawWalls = randomBool(rng);
        const ColorRgb wallColor = randomLayoutColor(rng);
        vg.addPlatform(*platform, layoutColor, wallColor, drawWalls);

        buildingZone = platform->terrainBoxes[TERRAIN_BUILDING_ZONE].front().boundingBox();
        currBuildingZoneReward = 0.0f;
        objectsInBuildingZone.clear();
        highestTower = 0;
        agentSpawnPositions = platform->agentSpawnCoords;
        objectSpawnPositions = platform->objectSpawnCoords;
        agentInitialPositions = agentSpawnPositions;
    }

    void spawnAgents() {  // scenario_default.hpp:80-97
        const float lookLimit = floatParams["verticalLookLimitRad"];
        const auto agentPositions = agentSpawnPositions;
        agents.resize(size_t(numAgents));
        for (int i = 0; i < numAgents; ++i) {
            auto randomRotation = frand(rng) * 3.14159265358979323846f * 2;
            agents[i].init(agentPositions[i] + Vec3{0.5f, 0.0f, 0.5f}, randomRotation, lookLimit);
            agents[i].updateTransform();
        }
    }

    // layout_utils.cpp:17-50 for every group of toBoundingBoxes (layout_utils.hpp:14-19), then terrain, then objects
    void towerAddEpisodeDrawables() {
        addDrawablesAndCollisionObjectsFromVoxelGrid(1.0f);
        for (auto &[terrainType, boxes] : platform->terrainBoxes)
            for (auto &bb : boxes) addTerrain(terrainType, bb.boundingBox());
        const auto objectPositions = objectSpawnPositions;
        for (const auto &pos : objectPositions)
            if (isInBuildingZone(pos)) objectsInBuildingZone.insert(pos);
        currBuildingZoneReward = calculateTowerReward();
        addObjects(objectPositions);
    }

    void addDrawablesAndCollisionObjectsFromVoxelGrid(float voxelSize) {  // layout_utils.hpp:12-18
        auto byType = vg.toBoundingBoxes();
        for (auto &[info, boxes] : byType) addBoundingBoxes(boxes, info.type, info.color, voxelSize);
    }
    void addBoundingBoxes(const Boxes &boxes, int voxelType, ColorRgb color, float voxelSize) {  // layout_utils.cpp:17-50
        if (voxelType == VOXEL_EMPTY) return;
        for (auto &box : boxes) {
            staticBoxes.push_back({box, uint8_t(voxelType), color});
            if (voxelType & VOXEL_OPAQUE)
                drawables[MESH_BOX].push_back({DrawEntry::D_STATIC, 0, paletteIndex(color),
                                               mul(mat4Translation(staticBoxTranslation(box, voxelSize)), mul(mat4Scaling(staticBoxScale(box, voxelSize)), mat4Identity()))});
            if (voxelType & VOXEL_SOLID) {
                Collider c;
                c.kind = 0;
                c.h = staticBoxScale(box, voxelSize);
                c.c = staticBoxTranslation(box, voxelSize);
                colliders.push_back(c);
            }
        }
    }
    static Vec3 staticBoxScale(const BoundingBox &b, float vs) {
        return Vec3{float(b.max.x - b.min.x + 1) / 2, float(b.max.y - b.min.y + 1) / 2, float(b.max.z - b.min.z + 1) / 2} * vs;
    }
    static Vec3 staticBoxTranslation(const BoundingBox &b, float vs) {
        return Vec3{float(b.min.x + b.max.x) / 2 + 0.5f, float(b.min.y + b.max.y) / 2 + 0.5f, float(b.min.z + b.max.z) / 2 + 0.5f} * vs;
    }

    void addTerrain(int terrain, const BoundingBox &bb) {  // layout_utils.cpp:53-68
        terrainSlabs.push_back({terrain, bb});
        const Vec3 scale = Vec3{float(bb.max.x - bb.min.x), 1.0f, float(bb.max.z - bb.min.z)} * 1.0f;
        if (scale.x > 0) {
            const Vec3 pos{bb.min.x * 1.0f + scale.x / 2, bb.min.y * 1.0f, bb.min.z * 1.0f + scale.z / 2};
            Mat4 m = mul(mat4Scaling({0.5f, 0.025f, 0.5f}), mat4Identity());
            m = mul(mat4Scaling(scale), m);
            m = mul(mat4Translation({0.0f, 0.025f, 0.0f}), m);
            m = mul(mat4Translation(pos), m);
            drawables[MESH_BOX].push_back({DrawEntry::D_STATIC, 0, paletteIndex(terrainColor(terrain)), m});
        }
    }
    void addRewardDrawables(const RewardObject &ro) {  // addDiamond: root, then its bottom half (layout_utils.cpp:122-123)
        const int idx = int(rewardObjects.size());
        drawables[MESH_CONE].push_back({DrawEntry::D_REWARD_ROOT, idx, ro.color, mat4Identity()});
        drawables[MESH_CONE].push_back({DrawEntry::D_REWARD_BOTTOM, idx, ro.color, mat4Identity()});
    }
    void addStaticCollidingBox(Vec3 scale, Vec3 translation, ColorRgb color) {  // layout_utils.cpp:70-84
        const Mat4 m = mul(mat4Translation(translation), mul(mat4Scaling(scale), mat4Identity()));
        drawables[MESH_BOX].push_back({DrawEntry::D_STATIC, 0, paletteIndex(color), m});
        Collider c;
        c.kind = 0;
        c.c = translationOf(m) + Vec3{0, 0, 0};
        c.h = scalingOf(m) * Vec3{1, 1, 1};
        colliders.push_back(c);
    }

    void addObjects(const std::vector<VoxelCoords> &positions) {  // component_object_stacking.hpp:170-198
        const float objSize = 0.39f;
        for (const auto &pos : positions) {
            MovableObject o;
            const Vec3 translation{float(pos.x) + 0.5f, float(pos.y) + 0.5f, float(pos.z) + 0.5f};
            o.local = mul(mat4Translation(translation), mul(mat4Scaling({objSize, objSize, objSize}), mat4Identity()));
            o.color = paletteIndex(MOVABLE_BOX);
            o.collider = int(colliders.size());
            colliders.push_back(Collider{});
            objects.push_back(o);
            drawables[MESH_BOX].push_back({DrawEntry::D_OBJECT, int(objects.size()) - 1, o.color, mat4Identity()});
            syncPose(int(objects.size()) - 1);
            if (!vg.grid.hasVoxel(pos)) vg.grid.set(pos, Voxel{});
            vg.grid.get(pos)->physicsObject = int(objects.size()) - 1;
        }
    }
    Mat4 objectAbs(int idx) const {  // Object::absoluteTransformation (left-to-right composition)
        const MovableObject &o = objects[idx];
        if (o.parentAgent < 0) return o.local;
        return mul(agents[o.parentAgent].pickupAbs(), o.local);
    }
    void syncPose(int idx) {  // physics.hpp:69-74
        const Mat4 m = objectAbs(idx);
        Collider &c = colliders[objects[idx].collider];
        c.kind = 0;
        c.c = translationOf(m) + objects[idx].collisionOffset;
        c.h = scalingOf(m) * objects[idx].collisionScale;
    }

    void addAgentsAndUI() {  // scenario_default.hpp:99-162 ; ghost objects join the collision world (agent.cpp:63)
        for (int i = 0; i < numAgents; ++i) {
            Agent &a = agents[i];
            a.color = paletteIndex(agentColors[i % numAgentColors]);
            a.bodyLocal = mul(mat4Translation({0, 0.09f, 0}), mul(mat4Scaling({0.35f, 0.36f, 0.35f}), mat4Identity()));
            a.eyesLocal = mul(mat4Translation({0.0f, 0.0f, -0.19f}), mul(mat4Scaling({0.25f, 0.12f, 0.2f}), mat4Identity()));
            a.uiLocal = mul(mat4Translation({0, 0, -0.2f}), mat4Identity());
            a.barAnchorLocal = mul(mat4Translation({0, -0.131f, 0}), mat4Identity());
            a.barLocal = mul(mat4Identity(), mat4Scaling({0.24f, float(0.0015), float(0.001)}));
        }
        for (int i = 0; i < numAgents; ++i) drawables[MESH_BOX].push_back({DrawEntry::D_EYES, i, paletteIndex(AGENT_EYES), mat4Identity()});
        for (int i = 0; i < numAgents; ++i) drawables[MESH_BOX].push_back({DrawEntry::D_BAR, i, paletteIndex(BLUE), mat4Identity()});
        for (int i = 0; i < numAgents; ++i) drawables[MESH_CAPSULE].push_back({DrawEntry::D_BODY, i, agents[i].color, mat4Identity()});
        agentColliderBase = int(colliders.size());
        for (int i = 0; i < numAgents; ++i) {
            Collider c;
            c.kind = 1;
            c.c = agents[i].kcc.pos;
            colliders.push_back(c);
        }
    }

    // ---------------------------------------------------------------- step (env.cpp:83-152)
    void setAction(int agentIdx, int mask) { currAction[agentIdx] = mask; }

    void step() {
        std::fill(lastReward.begin(), lastReward.end(), 0.0f);
        teleportLog.clear();
        const float dt = lastFrameDurationSec;
        for (int i = 0; i < numAgents; ++i) {
            const int a = currAction[i];
            Agent &agent = agents[i];
            Vec3 acceleration{0, 0, 0};
            if (a & A_Forward) acceleration += agent.forwardDirection();
            else if (a & A_Backward) acceleration -= agent.forwardDirection();
            if (a & A_Left) acceleration += agent.strafeLeftDirection();
            else if (a & A_Right) acceleration -= agent.strafeLeftDirection();
            if (a & A_LookLeft) agent.lookLeft(dt);
            else if (a & A_LookRight) agent.lookRight(dt);
            if (a & A_LookUp) agent.lookUp(dt);
            else if (a & A_LookDown) agent.lookDown(dt);
            agent.kcc.setAcceleration(acceleration, dt);
            if (a & A_Jump) agent.jump();
        }
        // stepSimulation(dt, 1, dt): action interfaces in insertion (agent) order (env.cpp:126)
        for (int i = 0; i < numAgents; ++i) {
            agents[i].kcc.playerStep(colliders, agentColliderBase + i, simulationStepSeconds);
            colliders[agentColliderBase + i].c = agents[i].kcc.pos;
        }
        for (auto &a : agents) a.updateTransform();

        if (scenario == S_TOWER) towerStep(); else if (scenario == S_OBSTACLES) obstaclesStep(); else if (scenario == S_COLLECT) collectStep();
        else if (scenario == S_REARRANGE) rearrangeStep(); else if (scenario == S_SOKOBAN) sokobanStep(); else if (scenario == S_HEX_EXPLORE) hexExploreStep();
        else if (scenario == S_EMPTY) {}  // EmptyScenario::step (scenario_empty.hpp:20)
        else hexMemoryStep();

        currEpisodeSec += lastFrameDurationSec;
        updateUI();
        if (currEpisodeSec >= episodeLengthSec()) done = true;
        for (int i = 0; i < numAgents; ++i) currAction[i] = 0;
        for (int i = 0; i < numAgents; ++i) totalReward[i] += lastReward[i];
        ++numFrames;
    }

    float episodeLengthSec() const {
        if (scenario == S_HEX_MEMORY) return floatParams.at("episodeLengthSec") + 3.0f * goodObjects.size();  // scenario_hex_memory.hpp:51-55
        if (scenario == S_REARRANGE || scenario == S_SOKOBAN || scenario == S_HEX_EXPLORE || scenario == S_EMPTY) return floatParams.at("episodeLengthSec");  // Scenario::episodeLengthSec (scenario.hpp:174-178)
        if (scenario == S_COLLECT) return floatParams.at("episodeLengthSec") + 2.0f * rewardSpawnPositions.size();  // scenario_collect.hpp:52-56
        if (scenario == S_OBSTACLES)  // scenario_obstacles.cpp:262-266
            return std::max(floatParams.at("episodeLengthSec"), float(numPlatforms) * 35 + float(objectSpawnPositions.size()) * 1);
        return floatParams.at("episodeLengthSec") + 4.0f * float(objectSpawnPositions.size());  // scenario_tower_building.cpp:263-266
    }
    float remainingTimeFraction() const { const float len = episodeLengthSec(); return std::max(0.0f, (len - currEpisodeSec) / len); }  // env.hpp:224-228
    float trueObjective(int) const { return scenario == S_TOWER ? float(highestTower) : (scenario == S_EMPTY ? 0.0f : float(solved)); }
    void doneWithTimer(float remaining = 0.3f) { currEpisodeSec = std::max(currEpisodeSec, episodeLengthSec() - remaining); }

    void updateUI() {  // scenario_default.hpp:164-186 ; UIElement::rescale :33-37
        for (int i = 0; i < numAgents; ++i) {
            Agent &a = agents[i];
            const Vec3 required{remainingTimeFraction() * 0.24f, float(0.0015), float(0.001)};
            const Vec3 scale = scalingOf(a.barLocal);
            a.barLocal = mul(a.barLocal, mat4Scaling({required.x / scale.x, required.y / scale.y, required.z / scale.z}));
        }
    }

    // reward plumbing (scenario.hpp:251-307)
    float getReward(const std::string &name, int agentIdx) const { return rewardShaping[agentIdx].at(name); }
    void rewardAgent(const std::string &name, int agentIdx, float multiplier) { lastReward[agentIdx] += getReward(name, agentIdx) * multiplier; }
    float teamSpirit(int agentIdx) const { return getReward("teamSpirit", agentIdx); }
    void rewardTeam(const std::string &name, int agentIdx, float multiplier) {
        rewardAgent(name, agentIdx, multiplier * (1 - teamSpirit(agentIdx)));
        for (int i = 0; i < numAgents; ++i) lastReward[i] += getReward(name, i) * teamSpirit(i) * multiplier / numAgents;
    }

    // ---------------------------------------------------------------- TowerBuilding step (scenario_tower_building.cpp:179-260)
    void towerStep() {
        for (int i = 0; i < numAgents; ++i)
            if (currAction[i] & A_Interact) onInteractAction(i);
        fallDetectionStep();
        for (int i = 0; i < numAgents; ++i) {
            if (carryingObject[i] >= 0) {
                const Vec3 t = translationOf(agents[i].objectT);
                VoxelCoords voxel = vg.grid.getCoords(t);
                if (isInBuildingZone(voxel)) {
                    if (!agentState[i].visitedBuildingZoneWithObject) {
                        rewardTeam("towerVisitedBuildingZoneWithObject", i, 1);
                        agentState[i].visitedBuildingZoneWithObject = true;
                    }
                }
            }
        }
    }
    bool isInBuildingZone(const VoxelCoords &c) const {
        return c.x >= buildingZone.min.x && c.x < buildingZone.max.x && c.z >= buildingZone.min.z && c.z < buildingZone.max.z;
    }
    static float buildingRewardCoeffForHeight(float height) {
        auto res = height * 0.05f;
        res += std::min(0.05f * powf(2, height), 20.0f);
        return res;
    }
    float calculateTowerReward() const {
        float reward = 0.0f;
        for (auto &pos : objectsInBuildingZone) reward += buildingRewardCoeffForHeight(float(pos.y));
        return reward;
    }
    void placedObject(int agentIdx, const VoxelCoords &voxel) {
        if (isInBuildingZone(voxel)) objectsInBuildingZone.insert(voxel);
        auto newReward = calculateTowerReward();
        auto rewardDelta = newReward - currBuildingZoneReward;
        currBuildingZoneReward = newReward;
        rewardTeam("towerBuildingReward", agentIdx, rewardDelta);
        highestTower = std::max(highestTower, voxel.y - buildingZone.min.y + 1);
    }
    void pickedObject(int agentIdx, const VoxelCoords &voxel) {
        if (isInBuildingZone(voxel)) objectsInBuildingZone.erase(voxel);
        if (!agentState[agentIdx].pickedUpObject) {
            rewardAgent("towerPickedUpObject", agentIdx, 1);
            agentState[agentIdx].pickedUpObject = true;
        }
    }

    void onInteractAction(int agentIdx) {  // component_object_stacking.hpp:58-168
        Agent &agent = agents[agentIdx];
        const float carryingScale = 0.78f, carryingScaleInverse = 1.0f / carryingScale;
        if (carryingObject[agentIdx] >= 0) {
            const int obj = carryingObject[agentIdx];
            const Vec3 t = translationOf(objectAbs(obj));
            VoxelCoords voxel = vg.grid.getCoords(t);
            auto voxelPtr = vg.grid.get(voxel);
            bool collidesWithAgent = false;
            for (int j = 0; j < numAgents; ++j) {
                if (j == agentIdx) continue;
                VoxelCoords c = vg.grid.getCoords(translationOf(agents[j].objectT));
                if (voxel == c) { collidesWithAgent = true; break; }
            }
            const bool empty = !voxelPtr || (voxelPtr->empty() && voxelPtr->physicsObject < 0);
            // canPlaceObject: TowerBuilding only inside the building zone (scenario_tower_building.cpp:201-204), else the default (true)
            bool canPlace = true;
            if (scenario == S_TOWER) canPlace = isInBuildingZone(voxel);
            else if (scenario == S_REARRANGE) canPlace = std::abs(voxel.x - rightCenter.x) <= 2 && std::abs(voxel.z - rightCenter.z) <= 2;  // scenario_rearrange.cpp:128-132
            if (empty && !collidesWithAgent && canPlace) {
                while (true) {
                    VoxelCoords below{voxel.x, voxel.y - 1, voxel.z};
                    if (below.y < -30) break;
                    auto belowPtr = vg.grid.get(below);
                    if (belowPtr && (belowPtr->solid() || belowPtr->physicsObject >= 0)) break;
                    else voxel = below;
                }
                if (!vg.grid.hasVoxel(voxel)) vg.grid.set(voxel, Voxel{});
                vg.grid.get(voxel)->physicsObject = obj;
                MovableObject &o = objects[obj];
                o.parentAgent = -1;
                const Vec3 scaling = scalingOf(o.local);
                o.local = mat4Identity();
                o.local = mul(mat4Scaling({scaling.x * carryingScaleInverse, scaling.y * carryingScaleInverse, scaling.z * carryingScaleInverse}), o.local);
                o.local = mul(mat4Translation({float(voxel.x) + 0.5f, float(voxel.y) + 0.5f, float(voxel.z) + 0.5f}), o.local);
                syncPose(obj);
                colliders[o.collider].enabled = !colliders[o.collider].enabled;  // toggleCollision
                carryingObject[agentIdx] = -1;
                if (scenario == S_TOWER) placedObject(agentIdx, voxel);
                else if (scenario == S_REARRANGE) { objects[obj].pickedUp = false; rearrangeCheckDone(agentIdx); }  // :151-155
            }
        } else {
            const Vec3 pickup = translationOf(agent.pickupAbs());
            VoxelCoords voxel = toVoxel(pickup);
            VoxelCoords voxelAbove{voxel.x, voxel.y + 1, voxel.z};
            int pickupHeight = 0, maxPickupHeight = 1;
            while (pickupHeight <= maxPickupHeight) {
                auto voxelPtr = vg.grid.get(voxel), voxelAbovePtr = vg.grid.get(voxelAbove);
                bool hasObjectAbove = voxelAbovePtr && voxelAbovePtr->physicsObject >= 0;
                if (voxelPtr && voxelPtr->physicsObject >= 0 && !hasObjectAbove) {
                    const int obj = voxelPtr->physicsObject;
                    MovableObject &o = objects[obj];
                    colliders[o.collider].enabled = !colliders[o.collider].enabled;
                    const Vec3 scaling = scalingOf(o.local);
                    o.local = mat4Identity();
                    o.local = mul(mat4Scaling({scaling.x * carryingScale, scaling.y * carryingScale, scaling.z * carryingScale}), o.local);
                    o.local = mul(mat4Translation({0.0f, -0.3f, 0.0f}), o.local);
                    o.parentAgent = agentIdx;
                    carryingObject[agentIdx] = obj;
                    voxelPtr->physicsObject = -1;
                    if (scenario == S_TOWER) pickedObject(agentIdx, voxel);
                    else if (scenario == S_REARRANGE) { objects[obj].pickedUp = true; rearrangeCheckDone(agentIdx); }  // :157-161
                    break;
                } else {
                    voxel = voxelAbove;
                    voxelAbove = VoxelCoords{voxel.x, voxel.y + 1, voxel.z};
                }
                ++pickupHeight;
            }
        }
    }

    void resetAgent(int i) {  // FallDetectionComponent::resetAgent, component_fall_detection.hpp:44-56
        Vec3 p = agentInitialPositions[i];
        auto v = vg.grid.getWithVector(p);
        while (v && !v->empty() && p.y < 1000) { p.y += 1; v = vg.grid.getWithVector(p); }
        const float halfVoxel = vg.grid.getVoxelSize() / 2;
        teleportLog.push_back({i, {p.x + halfVoxel, p.y + halfVoxel, p.z + halfVoxel}});
        agents[i].kcc.warp({p.x + halfVoxel, p.y + halfVoxel, p.z + halfVoxel});
        colliders[agentColliderBase + i].c = agents[i].kcc.pos;
    }
    void fallDetectionStep() {  // component_fall_detection.hpp:33-42
        for (int i = 0; i < numAgents; ++i)
            if (translationOf(agents[i].objectT).y < -20) resetAgent(i);
    }

    // ---------------------------------------------------------------- render interface
    // Instances in V4R draw order: mesh type major (meshIndices is a std::map<DrawableType,int>), insertion order minor
    // (v4r_env_renderer.cpp:267-279). Model matrices are the drawables' absoluteTransformationMatrix() (v4r_env_renderer.cpp:52-55),
    // which Magnum recomputes on every call as compose(parent.absoluteTransformation(), transformation()) (SceneGraph/Object.hpp:114-117):
    // the chain is multiplied LEFT to RIGHT from the root, ((agent * camera) * ui) * ...  -- pinned by tests/test_ref_shim.py against
    // the reference's own scene graph and scenario sources.
    std::vector<Instance> instances() const {
        std::vector<Instance> out;
        for (auto &[mesh, list] : drawables)
            for (auto &d : list) {
                Mat4 m = d.model;
                switch (d.kind) {
                    case DrawEntry::D_STATIC: break;
                    case DrawEntry::D_OBJECT: {
                        const MovableObject &o = objects[size_t(d.index)];
                        m = o.local;
                        if (o.parentAgent >= 0) {
                            const Agent &a = agents[size_t(o.parentAgent)];
                            m = mul(a.pickupAbs(), o.local);
                        }
                        break;
                    }
                    case DrawEntry::D_EYES: { const Agent &a = agents[size_t(d.index)]; m = mul(a.cameraAbs(), a.eyesLocal); break; }
                    case DrawEntry::D_BAR: { const Agent &a = agents[size_t(d.index)]; m = mul(mul(mul(a.cameraAbs(), a.uiLocal), a.barAnchorLocal), a.barLocal); break; }
                    case DrawEntry::D_BODY: { const Agent &a = agents[size_t(d.index)]; m = mul(a.objectT, a.bodyLocal); break; }
                    case DrawEntry::D_REWARD_ROOT: m = rewardObjects[size_t(d.index)].root; break;
                    case DrawEntry::D_REWARD_BOTTOM: m = mul(rewardObjects[size_t(d.index)].root, rewardObjects[size_t(d.index)].bottomLocal); break;
                    case DrawEntry::D_MEMORY: {
                        const MemoryObject &mo = memoryObjects[size_t(d.index / 4)];
                        const int child = d.index % 4;
                        m = child == 0 ? mo.root : mul(mo.root, mo.childLocal[child - 1]);
                        break;
                    }
                }
                out.push_back({mesh, d.color, m});
            }
        return out;
    }
    Mat4 viewMatrix(int agentIdx) const { return inverted(agents[agentIdx].cameraAbs()); }  // Camera::cameraMatrix

public:
    int scenario = S_TOWER;
    int numAgents;
    FloatParams floatParams;
    std::vector<RewardShaping> rewardShaping;
    Rng rng{std::random_device{}()};

    bool done = false;
    int numFrames = 0;
    float currEpisodeSec = 0;
    float simulationStepSeconds = 1.0f / 15.0f, lastFrameDurationSec = 1.0f / 15.0f;
    std::vector<int> currAction;
    std::vector<float> lastReward, totalReward;

    std::vector<Agent> agents;
    std::vector<std::pair<int, Vec3>> teleportLog;  // (agent, target) of this tick's AbstractAgent::teleport calls
    bool undefinedSpawn = false;                     // see obstaclesReset: the reference read past its spawn-point vector
    std::vector<Collider> colliders;
    int agentColliderBase = 0;
    std::vector<MovableObject> objects;
    std::vector<StaticBox> staticBoxes;
    std::vector<TerrainSlab> terrainSlabs;
    std::map<int, std::vector<DrawEntry>> drawables;
    // hexagonal mazes
    unsigned episodeSeed = 0;
    std::unique_ptr<HoneyCombMaze> hexMaze;
    int hexMazeSize = 0;
    float hexMazeScale = 1.0f, hexWallHeight = 1.0f, hexOmitWallsProbability = 0.0f, hexWallLandmarkProbability = 0.0f;
    ColorRgb hexBottomEdgingColor = WHITE, hexTopEdgingColor = WHITE;
    double hexXMin = 0, hexXMax = 0, hexYMin = 0, hexYMax = 0;
    Vec3 rewardObjectCoords{0, 0, 0};
    bool exploreRewardAlive = false;
    std::vector<Vec3> goodObjects, badObjects;
    std::vector<MemoryObject> memoryObjects;
    Vec3 landmarkLocation{0, 0, 0};
    int goodObjectsCollected = 0;
    // Sokoban
    std::vector<std::string> allSokobanLevelFiles;
    std::vector<SokobanLevel> sokobanLevels;
    SokobanLevel currLevel;
    int sokoLength = 0, sokoWidth = 0, numBoxes = 0, numBoxesOnGoal = 0;
    std::vector<Vec3> sokobanAgentPositions;
    std::vector<VoxelCoords> boxesCoords;
    // Rearrange
    std::unique_ptr<RearrangePlatform> rearrangePlatform;
    std::vector<ArrangementItem> arrangement;
    std::vector<int> arrangementObjects;
    int maxMatchingObjects = 0;
    const VoxelCoords leftCenter{5, 2, 5}, rightCenter{13, 2, 5};

    VoxelGridComponent vg;
    std::vector<int> carryingObject;
    std::vector<Vec3> agentInitialPositions;
    std::unique_ptr<Node> levelRoot;
    std::unique_ptr<TowerPlatform> platform;

    // generic level products
    std::vector<Vec3> agentSpawnPositions;
    std::vector<VoxelCoords> objectSpawnPositions, rewardSpawnPositions;
    // Obstacles
    std::vector<int> platformTypes;
    bool onePlatformType = false;
    std::vector<std::unique_ptr<Platform>> platforms;
    std::vector<bool> agentReachedExit;
    bool solved = false;
    int numPlatforms = 0;
    int numPositiveRewards = 0, positiveRewardsCollected = 0;  // Collect
    std::vector<RewardObject> rewardObjects;

    struct TowerAgentState { bool pickedUpObject = false, visitedBuildingZoneWithObject = false; };
    std::vector<TowerAgentState> agentState;
    int highestTower = 0;
    BoundingBox buildingZone;
    std::unordered_set<VoxelCoords, VoxelHash> objectsInBuildingZone;  // persists across episodes (clear() keeps the buckets)
    float currBuildingZoneReward = 0.0f;
};

}  // namespace orc
