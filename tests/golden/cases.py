"""Golden cases shared by the generator (gen_golden.py, needs /root/reference) and the tests."""
from romis_b200 import abi
from romis_b200.scene import Camera, Features, RmisParams

NIGHTCLUB_CAM = Camera()                                            # reference src/utils/config.h:21-26
CORNELL_CAM = Camera(50.0, 3.0, (0.0, 0.0, 0.0), (20.0, 20.0, 0.0))   # TOML fallback camera, src/utils/config.cpp:249-252

# name -> (scene, W, H, Features, Camera, frames, seed)
CASES = {
    "nightclub_default": ("CornellNightClub", 48, 32, Features(), NIGHTCLUB_CAM, 2, 7),
    "nightclub_c2": ("CornellNightClub", 40, 30, Features(spatialResamplingPasses=3, initialSamplesVisibilityCheck=True), NIGHTCLUB_CAM, 3, 11),
    "nightclub_unbiased_vis_n1": ("CornellNightClub", 32, 32, Features(unbiasedCombination=True, spatialReuseVisibilityCheck=True,
                                                                        numSamplesInReservoir=1), NIGHTCLUB_CAM, 2, 13),
    "nightclub_unbiased_n3": ("CornellNightClub", 32, 24, Features(unbiasedCombination=True, numSamplesInReservoir=3,
                                                                    numNeighboursToSample=3, spatialResampleRadius=4), NIGHTCLUB_CAM, 2, 17),
    "nightclub_n5_generic": ("CornellNightClub", 24, 24, Features(numSamplesInReservoir=5, initialLightSamples=16, temporalClampM=2), NIGHTCLUB_CAM, 3, 19),
    "cornell_c1": ("CornellBoxParallelogramLight", 48, 48, Features(spatialResamplingPasses=1), CORNELL_CAM, 2, 23),
    "cornell_point_n1": ("CornellBox", 40, 40, Features(spatialResamplingPasses=1, numSamplesInReservoir=1), CORNELL_CAM, 2, 29),
    "monkey": ("Monkey", 40, 40, Features(), CORNELL_CAM, 2, 31),
    "cube_segment": ("Cube", 32, 32, Features(numSamplesInReservoir=4, gamma=2.2, exposure=0.8), CORNELL_CAM, 2, 37),
    "cube_textured": ("CubeTextured", 32, 32, Features(), CORNELL_CAM, 2, 41),
    "triangle_noshading_k0": ("SingleTriangle", 24, 24, Features(enableShading=False, numNeighboursToSample=0, enableToneMapping=False), CORNELL_CAM, 2, 43),
    "nightclub_no_spatial": ("CornellNightClub", 32, 24, Features(spatialReuse=False), NIGHTCLUB_CAM, 3, 47),
    "nightclub_no_temporal": ("CornellNightClub", 32, 24, Features(temporalReuse=False, spatialResampleRadius=30, numNeighboursToSample=10), NIGHTCLUB_CAM, 2, 53),
}
SCENES = ["SingleTriangle", "Cube", "CubeTextured", "CornellBox", "CornellBoxParallelogramLight", "CornellNightClub", "Monkey"]

# R-MIS mode (renderRMIS, reference src/rendering/render.cpp:64-119): name -> (scene, W, H, Features, RmisParams, Camera, seed, frame)
RMIS_CASES = {
    "rmis_nightclub_similar_equal": ("CornellNightClub", 40, 30, Features(), RmisParams(maxIterationsMIS=3), NIGHTCLUB_CAM, 61, 0),
    "rmis_nightclub_similar_balance": ("CornellNightClub", 36, 28, Features(numSamplesInReservoir=1),
                                       RmisParams(maxIterationsMIS=2, misWeightRMIS=abi.ROMIS_MIS_BALANCE), NIGHTCLUB_CAM, 67, 3),
    "rmis_nightclub_random_balance": ("CornellNightClub", 32, 24, Features(numNeighboursToSample=3, spatialResampleRadius=30),
                                      RmisParams(maxIterationsMIS=2, misWeightRMIS=abi.ROMIS_MIS_BALANCE,
                                                 neighbourSelectionStrategy=abi.ROMIS_NEIGHBOURS_RANDOM), NIGHTCLUB_CAM, 71, 1),
    "rmis_monkey_esd_balance": ("Monkey", 40, 33, Features(numNeighboursToSample=10, spatialResampleRadius=4, numSamplesInReservoir=3),
                                RmisParams(maxIterationsMIS=2, misWeightRMIS=abi.ROMIS_MIS_BALANCE, neighbourSameGeometry=False,
                                           neighbourSelectionStrategy=abi.ROMIS_NEIGHBOURS_EQUAL_SIMILAR_DISSIMILAR), CORNELL_CAM, 73, 0),
    # window smaller than k: std::sample returns fewer than k neighbours
    "rmis_cube_small_window": ("CubeTextured", 24, 24, Features(numNeighboursToSample=10, spatialResampleRadius=1, initialSamplesVisibilityCheck=True,
                                                                 gamma=2.2, exposure=0.8),
                               RmisParams(maxIterationsMIS=1, neighbourMaxDepthDifferenceFraction=0.02), CORNELL_CAM, 79, 2),
    "rmis_cornell_esd_equal_r30": ("CornellBoxParallelogramLight", 40, 40, Features(spatialResampleRadius=30, numSamplesInReservoir=5, initialLightSamples=8),
                                   RmisParams(maxIterationsMIS=2, neighbourSelectionStrategy=abi.ROMIS_NEIGHBOURS_EQUAL_SIMILAR_DISSIMILAR), CORNELL_CAM, 83, 0),
}

# R-OMIS mode (renderROMIS, reference src/rendering/render.cpp:121-265), direct estimator: same tuple layout as RMIS_CASES
ROMIS_CASES = {
    "romis_nightclub_similar": ("CornellNightClub", 36, 27, Features(), RmisParams(maxIterationsMIS=3), NIGHTCLUB_CAM, 89, 0),
    "romis_nightclub_random_n1": ("CornellNightClub", 32, 24, Features(numSamplesInReservoir=1, numNeighboursToSample=3, spatialResampleRadius=30),
                                  RmisParams(maxIterationsMIS=2, neighbourSelectionStrategy=abi.ROMIS_NEIGHBOURS_RANDOM), NIGHTCLUB_CAM, 97, 2),
    "romis_monkey_esd_k7_n3": ("Monkey", 36, 30, Features(numNeighboursToSample=7, spatialResampleRadius=4, numSamplesInReservoir=3),
                               RmisParams(maxIterationsMIS=2, neighbourSameGeometry=False,
                                          neighbourSelectionStrategy=abi.ROMIS_NEIGHBOURS_EQUAL_SIMILAR_DISSIMILAR), CORNELL_CAM, 101, 0),
    "romis_cornell_vis_k10_n5": ("CornellBoxParallelogramLight", 28, 28, Features(numNeighboursToSample=10, spatialResampleRadius=6, numSamplesInReservoir=5,
                                                                                  initialLightSamples=8, initialSamplesVisibilityCheck=True, gamma=2.2),
                                 RmisParams(maxIterationsMIS=2), CORNELL_CAM, 103, 1),
    # progressive estimator (useProgressiveROMIS, render.cpp:133-139,160-200); N >= k + 1, or its integer N / (k + 1) is 0
    "romis_progressive_nightclub_n3_k2": ("CornellNightClub", 36, 27, Features(numSamplesInReservoir=3, numNeighboursToSample=2, spatialResampleRadius=4),
                                          RmisParams(maxIterationsMIS=3, useProgressiveROMIS=True), NIGHTCLUB_CAM, 109, 0),
    "romis_progressive_cornell_n6_k5_mod2": ("CornellBoxParallelogramLight", 28, 28, Features(numSamplesInReservoir=6, initialLightSamples=16),
                                             RmisParams(maxIterationsMIS=5, useProgressiveROMIS=True, progressiveUpdateMod=2,
                                                        neighbourSelectionStrategy=abi.ROMIS_NEIGHBOURS_RANDOM), CORNELL_CAM, 113, 1),
    "romis_cube_textured_notonemap": ("CubeTextured", 24, 24, Features(numNeighboursToSample=2, spatialResampleRadius=2, enableToneMapping=False),
                                      RmisParams(maxIterationsMIS=4, neighbourMaxDepthDifferenceFraction=0.02), CORNELL_CAM, 107, 0),
}
