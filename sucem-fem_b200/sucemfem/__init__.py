"""sucemfem -- B200-native drop-in for the assembly path of SUCEM:FEM."""
