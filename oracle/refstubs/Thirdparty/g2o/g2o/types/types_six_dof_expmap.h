// TEST INFRASTRUCTURE ONLY.  include/Converter.h names these g2o types in declarations; nothing on the compared path calls them.
#pragma once
namespace g2o {
class SE3Quat;
class Sim3;
}  // namespace g2o
