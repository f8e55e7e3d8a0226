// no optional extensions enabled (stand-in for the cmake-generated file)
#pragma once
