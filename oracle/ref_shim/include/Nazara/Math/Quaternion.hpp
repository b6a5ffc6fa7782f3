// oracle/ref_shim — stand-in for <Nazara/Math/Quaternion.hpp>: Quaternion lives with Vector3 in this shim.
#pragma once
#include <Nazara/Math/Vector3.hpp>
