// stand-in for include/MapPoint.h (members used by the adapters: :58 GetWorldPos, :67 isBad, :81 GetDescriptor, :70 Observations)
#pragma once
#include <opencv2/core/core.hpp>
namespace StructureSLAM { class MapPoint { public: cv::Mat GetWorldPos(); bool isBad(); cv::Mat GetDescriptor(); int Observations(); }; }
