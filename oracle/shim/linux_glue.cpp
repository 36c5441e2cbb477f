// oracle glue: gTimerPeriod is only defined under WIN32 in Utils/Timer.cpp but referenced by Utils/Profiler.cpp
namespace rt { extern const double gTimerPeriod; const double gTimerPeriod = 1.0e-9; }
