#pragma once
#include <fmt/format.h>
