"""meshdiffusion_b200: B200-native (sm_100a) implementation of MeshDiffusion's score-network / sampler / marching-tet hot path."""
__version__ = "0.1.0"
