// TEST INFRASTRUCTURE. Stand-in for the header cmake generates from
// MGARDConfig.hpp.in (versions: reference CMakeLists.txt:13-19).
#pragma once
#define MGARD_VERSION_MAJOR 1
#define MGARD_VERSION_MINOR 6
#define MGARD_VERSION_PATCH 0
#define MGARD_FILE_VERSION_MAJOR 1
#define MGARD_FILE_VERSION_MINOR 0
#define MGARD_FILE_VERSION_PATCH 0
