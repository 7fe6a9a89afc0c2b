"""Host build of the device solver source (TEST INFRASTRUCTURE, see hostsim.cpp)."""
