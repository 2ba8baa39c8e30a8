// Runs the reference's own host-compiled OptiXRendererTests (BSDFs, shading models, lights,
// misc) from the staged copy in baseline/_ref. TEST INFRASTRUCTURE: this is the gate that
// pins the oracle build (same headers, same flags) to the reference's golden vectors.
// Mirrors tests/OptiXRendererTests/main.cpp:9-35 minus RendererTest.h (needs OptiX + GPU),
// LTCTest.h and BurleySSSTest.h (not on the hot path).
#include <gtest/gtest.h>

#include <Assets/FlagsTest.h>

#include <BSDFs/BurleyTest.h>
#include <BSDFs/GGXTest.h>
#include <BSDFs/LambertTest.h>
#include <BSDFs/OrenNayarTest.h>

#include <LightSources/SphereLightTest.h>
#include <LightSources/SpotLightTest.h>

#include <ShadingModels/DefaultShadingTest.h>
#include <ShadingModels/TransmissiveShadingTest.h>
#include <ShadingModels/UtilsTest.h>

#include <MiscTest.h>

int main(int argc, char** argv) {
    testing::InitGoogleTest(&argc, argv);
    return RUN_ALL_TESTS();
}
