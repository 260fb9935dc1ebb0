// ORACLE tooling.  Stand-in for the reference's scenarios/init.hpp: the same registration calls (init.hpp:28-56) for the scenarios this
// repository covers plus the debugging Empty / Test ones, without the experimental Football / BoxAGone, whose sources create dynamic Bullet
// bodies the stand-in Bullet (ref_shim/mini_bullet) does not model.
#pragma once
#include <env/scenario.hpp>

#include <scenarios/scenario_collect.hpp>
#include <scenarios/scenario_empty.hpp>
#include <scenarios/scenario_hex_explore.hpp>
#include <scenarios/scenario_hex_memory.hpp>
#include <scenarios/scenario_obstacles.hpp>
#include <scenarios/scenario_rearrange.hpp>
#include <scenarios/scenario_sokoban.hpp>
#include <scenarios/scenario_tower_building.hpp>

namespace Megaverse {

template <typename ScenarioType> void registerScenario(const std::string &name) { Scenario::registerScenario(name, Scenario::scenarioFactory<ScenarioType>); }

inline void scenariosGlobalInit() {
    static bool initialized = false;
    if (initialized) return;
    initialized = true;
    registerScenario<EmptyScenario>("Empty");
    registerScenario<TestScenario>("Test");
    registerScenario<TowerBuildingScenario>("TowerBuilding");
    registerScenario<ObstaclesEasyScenario>("ObstaclesEasy");
    registerScenario<ObstaclesHardScenario>("ObstaclesHard");
    registerScenario<CollectScenario>("Collect");
    registerScenario<SokobanScenario>("Sokoban");
    registerScenario<HexMemoryScenario>("HexMemory");
    registerScenario<HexExploreScenario>("HexExplore");
    registerScenario<RearrangeScenario>("Rearrange");
    registerScenario<ObstaclesMediumScenario>("ObstaclesMedium");
    registerScenario<ObstaclesOnlyWallsScenario>("ObstaclesWalls");
    registerScenario<ObstaclesOnlyStepsScenario>("ObstaclesSteps");
    registerScenario<ObstaclesOnlyLavaScenario>("ObstaclesLava");
}

}  // namespace Megaverse
