// TEST INFRASTRUCTURE (oracle/shim): see ros/ros.h
#pragma once
#include <ros/ros.h>
