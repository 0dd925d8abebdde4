"""B200-native ERI + Fock-build engine behind Quiqbox.jl's API (host-side mirror)."""
from .basis import (GTO, MultiOrbitalData, NuclearCluster, SubshellXYZs, genGaussTypeOrb,
                    genGaussTypeOrbSeq, get3DimPGTOrbNormFactor, nucRepulsion)
from .hartreefock import (HFconfig, HFfinalInfo, RCHartreeFock, SCFconfig, UOHartreeFock,
                          runHartreeFockCore)
from .hartreefock import runHartreeFock
from .integrals import (DeviceBasis, DeviceERI, DeviceSCF, boys, changeOrbitalBasis, coreHamiltonian, elecKinetics, elecRepulsion,
                        elecRepulsionList, elecRepulsions, getGcore, nucAttractions, overlaps)
