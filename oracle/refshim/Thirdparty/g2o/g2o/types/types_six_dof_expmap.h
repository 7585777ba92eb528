// TEST INFRASTRUCTURE ONLY: include/Converter.h names these g2o types in declarations; the hot path never touches them.
#pragma once
#include <Eigen/Core>
namespace g2o { class SE3Quat; class Sim3; }
