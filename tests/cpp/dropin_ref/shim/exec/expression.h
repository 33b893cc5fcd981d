#pragma once
#include "bosql_operator.hpp"
