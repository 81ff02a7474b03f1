"""Scene text of the synthetic benchmark configurations C4 / C5:
a height-field trianglemesh under the Cornell box's light, side walls and back wall, seen by
the Cornell camera.  The mesh itself is generated inside the host library by the
``heightfield_lcg`` shape extension (same vertices / indices a ``trianglemesh`` would list)."""


def heightfield_scene(nx, nz, xres, yres, spp, seed=20260418, integrator="path", max_depth=5):
    quad = lambda p: 'Shape "trianglemesh" "integer indices" [0 1 2 2 3 0] "point P" [%s]\n' % " ".join("%g" % v for v in p)
    return (
        'LookAt 278 273 -800 278 273 0 0 1 0\n'
        'Camera "perspective" "float fov" [37.5]\n'
        'Sampler "halton" "integer pixelsamples" [%d]\n'
        'Film "image" "string filename" ["heightfield.png"] "integer xresolution" [%d] "integer yresolution" [%d]\n'
        'Integrator "%s" "integer maxdepth" [%d]\n'
        'WorldBegin\n'
        'AttributeBegin\nAreaLightSource "diffuse" "rgb L" [17 12 4]\n%sAttributeEnd\n'
        'Material "matte" "rgb Kd" [0.582553 0.043657 0.044207]\n%s'
        'Material "matte" "rgb Kd" [0.100772 0.376082 0.078298]\n%s'
        'Material "matte" "rgb Kd" [0.886975 0.696752 0.667831]\n%s'
        'Material "matte" "rgb Kd" [0.5 0.5 0.5]\n'
        'Shape "heightfield_lcg" "integer nx" [%d] "integer nz" [%d] "integer seed" [%d]\n'
        'WorldEnd\n' % (
            spp, xres, yres, integrator, max_depth,
            quad([343, 548.7, 227, 343, 548.7, 332, 213, 548.7, 332, 213, 548.7, 227]),
            quad([552.8, 0, 0, 549.6, 0, 559.2, 556, 548.8, 559.2, 556, 548.8, 0]),
            quad([0, 0, 559.2, 0, 0, 0, 0, 548.8, 0, 0, 548.8, 559.2]),
            quad([549.6, 0, 559.2, 0, 0, 559.2, 0, 548.8, 559.2, 556, 548.8, 559.2]),
            nx, nz, seed))


CONFIGS = {
    "C4": dict(nx=1000, nz=500, xres=1920, yres=1080, spp=256),     # 1,000,000 triangles
    "C5": dict(nx=2500, nz=2000, xres=3840, yres=2160, spp=4096),   # 10,000,000 triangles
}
