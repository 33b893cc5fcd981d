#pragma once
// include shim: the reference spells this header differently; the product keeps these declarations in one file
#include "bosql_types.hpp"
