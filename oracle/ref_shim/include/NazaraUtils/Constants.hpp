// oracle/ref_shim — stand-in for <NazaraUtils/Constants.hpp>
#pragma once
#include <NazaraUtils/Prerequisites.hpp>
