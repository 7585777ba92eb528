// stand-in for include/MapLine.h (isBad)
#pragma once
#include <opencv2/core/core.hpp>
namespace StructureSLAM { class MapLine { public: bool isBad(); }; }
