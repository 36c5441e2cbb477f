// oracle stub of Demo/Demo.h: only what Demo/SceneLoader.cpp and Demo/MeshLoader.cpp need (no window, no xcb)
#pragma once
#include <string>
#include <map>
#include "../Core/Scene/Scene.h"
#include "../Core/Scene/Camera.h"
#include "../Core/Rendering/Viewport.h"
#include "../Core/Utils/Bitmap.h"
#include "../Core/Rendering/Context.h"
struct Options
{
    uint32 windowWidth = 1280;
    uint32 windowHeight = 720;
    std::string dataPath;
    uint32 numThreads = 0;
    bool enablePacketTracing = false;
    std::string rendererName = "Path Tracer";
    std::string sceneName;
};
extern Options gOptions;
