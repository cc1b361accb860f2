"""Python-3 restatement of the standard part of the reference's ``make_node_selections.py``
(:28-85): one node per residue ('COM') or per atom ('ATOMIC'), source / sink lookup by
``RESNAME_RESID`` strings, ``node_selections.txt``.  Boundary code, O(residues), runs once;
the hand-made nucleic-acid / ATP sub-residue nodes (:121-388) are outside the accelerated
path -- use the reference's own module for those (any list of AtomGroups works downstream:
the node map is a general CSR gather)."""
import sys


def make_selections(analysis_universe, file_name, node_definition, selection_string,
                    source_selection_strings, sink_selection_strings,
                    nonstandard_substrates_selection=None, homemade_selections=None):
    selection_list = []
    count = 0
    source_index_list = []
    sink_index_list = []
    with open(file_name, 'w') as f:
        if node_definition.upper() == 'COM':                                   # :39-58
            substrate_selection = analysis_universe.select_atoms(selection_string)
            for i in range(substrate_selection.n_residues):
                temp = substrate_selection.residues[i].atoms
                selection_list.append(temp)
                temp_resname = temp.resnames[0]
                temp_resid = temp.resids[0]
                node_string = '%s_%s' % (temp_resname, temp_resid)
                if node_string in source_selection_strings:
                    f.write("%02d   %s   %s # SOURCE\n" % (count, temp_resname, temp_resid))
                    source_index_list.append(count)
                elif node_string in sink_selection_strings:
                    f.write("%02d   %s   %s # SINK\n" % (count, temp_resname, temp_resid))
                    sink_index_list.append(count)
                else:
                    f.write("%02d   %s   %s\n" % (count, temp_resname, temp_resid))
                count += 1
        elif node_definition.upper() == 'ATOMIC':                              # :63-80
            substrate_selection = analysis_universe.select_atoms(selection_string)
            for i in range(substrate_selection.n_atoms):
                temp = substrate_selection.atoms[i]
                selection_list.append(temp)
                temp_resname = temp.resname
                temp_resid = temp.resid
                node_string = '%s_%s' % (temp_resname, temp_resid)
                if node_string in source_selection_strings:
                    f.write("%02d   %s   %s # SOURCE\n" % (count, temp_resname, temp_resid))
                    source_index_list.append(count)
                elif node_string in sink_selection_strings:
                    f.write("%02d   %s   %s # SINK\n" % (count, temp_resname, temp_resid))
                    sink_index_list.append(count)
                else:
                    f.write("%02d   %s   %s\n" % (count, temp_resname, temp_resid))
                count += 1
        else:
            print('The substrate_node_definition parameter is not understood.')
            sys.exit()
        if nonstandard_substrates_selection is not None:
            print('Hand-made nonstandard substrate nodes are not part of the B200 hot path; '
                  'use the reference make_node_selections.py for them.')
            sys.exit()
    return selection_list, source_index_list, sink_index_list
