#pragma once
#include "bosql_sql.hpp"
